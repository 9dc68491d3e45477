"""`fn_timer` of the reference (auxiliary_funs.py:22-30): wall-clock print around the driver."""
import functools
import time


def fn_timer(function):
    @functools.wraps(function)
    def wrapper(*args, **kwargs):
        t0 = time.time()
        out = function(*args, **kwargs)
        print(f"Total time running {function.__name__}: {time.time() - t0}")
        return out
    return wrapper
