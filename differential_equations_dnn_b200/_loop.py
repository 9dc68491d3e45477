"""CUDA-graph replay of a training iteration, shared by the four drivers.

At the reference's batch sizes (32-256 rows) a step is launch-bound: ~150 kernel launches plus the
Python between them.  `graphed_loop` runs the first iterations eagerly (they also size the caches
the capture relies on), captures ONE iteration -- device-side sampler, fused loss + gradient, fused
Adam with a device-resident step counter (`FusedAdam(capturable=True)`), loss record -- and replays
it.  Same RNG stream and arithmetic as the eager loop; losses stay on the device until the end.
"""
import torch

# timing of the last graphed_loop call: {"replays": n, "ms": CUDA-event time of the n replays} (bench.py's driver latency)
last_timing = {"replays": 0, "ms": 0.0}


def graphed_loop(step, iterations, device, warmup=11, counter=None):
    """step() -> 0-dim loss tensor; it must do its own zero_grad / backward / optimizer.step with
    device-side state only.  Returns the list of losses (one host read at the end).
    counter: a zeroed 1-element int64 device tensor to use as the iteration counter -- the on-device sampler
    (`sampler.PhiloxSampler.step`) reads it, so every replay draws fresh points with no extra launch."""
    losses = torch.zeros(max(iterations, 1), device=device)
    idx = counter if counter is not None else torch.zeros(1, dtype=torch.int64, device=device)

    def one_iteration():
        loss = step()
        losses.index_copy_(0, idx, loss.detach().reshape(1))
        idx.add_(1)

    n_eager = min(warmup, iterations)
    side = torch.cuda.Stream(device=device)
    side.wait_stream(torch.cuda.current_stream(device))
    with torch.cuda.stream(side):          # warm-up on the stream family the capture will use
        for _ in range(n_eager):
            one_iteration()
    torch.cuda.current_stream(device).wait_stream(side)
    if iterations > n_eager:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            one_iteration()
        # the capture itself does not execute: iteration n_eager is the first replay
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iterations - n_eager):
            graph.replay()
        e1.record()
        out = losses[:iterations].cpu().tolist()   # synchronises
        last_timing.update(replays=iterations - n_eager, ms=e0.elapsed_time(e1))
        return out
    last_timing.update(replays=0, ms=0.0)
    return losses[:iterations].cpu().tolist()


def print_progress(train_loss, lrate, rank):
    if rank == 0:
        for i in range(0, len(train_loss), 100):
            print(f"Iteration: {i}, Loss: {train_loss[i]}, LR: {lrate}")
