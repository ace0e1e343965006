"""Worker of tests/test_gpu_peer_exchange.py: run under torch.distributed.run, one process per GPU.
Checks rhseg_xchg_all_reduce (one-shot all-reduce over peer memory) against the NCCL all-reduce of the same
buffer: eager, with changing sizes, rank-skewed arrival, and replayed inside a CUDA graph."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rhseg_b200 import dist as rdist  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    px = rdist.PeerExchange(capacity=40000)

    def check(n_sum, part_sizes, tag):
        summary = torch.randn(n_sum, generator=g, dtype=torch.float64).to(dev)
        parts = [torch.randn(s, generator=g).to(dev) for s in part_sizes]
        want = rdist.pack_exchange(summary, parts)
        dist.all_reduce(want)
        got = px.all_reduce(summary, parts)
        torch.cuda.synchronize()
        assert got.shape == want.shape, (tag, got.shape, want.shape)
        err = (got - want).abs().max().item() if got.numel() else 0.0
        assert err <= 1e-12 * max(1.0, want.abs().max().item()), (tag, err)
        # bit-identical on every rank (fixed summation order)
        ref = got.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(ref, got), tag

    for it, (n_sum, sizes) in enumerate([(76, [2880, 2880, 4, 4, 5760, 1440]), (1, []), (10, [1]), (0, [33000]),
                                         (5000, [7, 0, 12000, 3]), (76, [2880, 2880, 4, 4, 5760, 1440])] * 3):
        if (it + rank) % 3 == 0:
            time.sleep(0.02)  # skewed arrival: peers spin on the flags meanwhile
        check(n_sum, sizes, "eager%d" % it)

    # in-place on a [summary | tail] buffer, replayed from a CUDA graph with fresh inputs per replay
    n_sum, sizes = 76, [2880, 2880, 4, 4, 5760, 1440]
    xbuf = torch.zeros(n_sum + sum(sizes), dtype=torch.float64, device=dev)
    summary = xbuf[:n_sum]
    parts = [torch.zeros(s, device=dev) for s in sizes]
    src_sum = torch.zeros(n_sum, dtype=torch.float64, device=dev)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            summary.copy_(src_sum)
            px.all_reduce(summary, parts, out=xbuf)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        summary.copy_(src_sum)  # stands for rhseg_step_finalize writing the summary each step
        out = px.all_reduce(summary, parts, out=xbuf)
    for it in range(25):
        src_sum.copy_(torch.randn(n_sum, generator=g, dtype=torch.float64))
        for p in parts:
            p.copy_(torch.randn(p.shape, generator=g))
        want = rdist.pack_exchange(src_sum, parts)
        dist.all_reduce(want)
        graph.replay()
        torch.cuda.synchronize()
        err = (out - want).abs().max().item()
        assert err <= 1e-12 * max(1.0, want.abs().max().item()), ("graph%d" % it, err)
    assert px.status() == 0
    # timing: peer-memory kernel vs rhseg_pack_f64 + NCCL, 200 back-to-back launches each
    def say(msg):
        if rank == 0:
            print(msg, flush=True)

    def timed(fn, n=200):
        for _ in range(5):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    def graphed(fn, reps=10):
        """`reps` back-to-back exchanges captured in one graph: device time per exchange without host launch cost."""
        s2 = torch.cuda.Stream()
        s2.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s2):
            fn()
        torch.cuda.current_stream().wait_stream(s2)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(reps):
                fn()
        return lambda: gr.replay(), reps
    t_p2p = timed(lambda: px.all_reduce(summary, parts, out=xbuf))
    t_nccl = timed(lambda: dist.all_reduce(rdist.pack_exchange(summary, parts, out=xbuf)))
    say("eager timing done: p2p %.1f us nccl %.1f us" % (t_p2p, t_nccl))
    f, reps = graphed(lambda: px.all_reduce(summary, parts, out=xbuf))
    say("p2p graph captured")
    t_p2p_g = timed(f, 40) / reps
    say("p2p graph timed %.2f us; status %d" % (t_p2p_g, px.status()))
    say("in-graph per exchange: p2p %.2f us" % t_p2p_g)
    del f, graph
    px.close()
    if rank == 0:
        print("PEER_EXCHANGE_OK world=%d p2p_us=%.1f nccl_us=%.1f" % (world, t_p2p, t_nccl), flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
