"""One-shot all-reduce over NVLink peer memory (rhseg_xchg_*): needs >= 2 GPUs on the box; the single-GPU
round-end run skips it.  The worker compares against NCCL, eager and inside a CUDA graph."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_memory_all_reduce_matches_nccl():
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(HERE, "peer_exchange_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "PEER_EXCHANGE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_peer_exchange_argument_errors():
    """Single-GPU checks of the C-ABI argument handling (no peers involved)."""
    import ctypes
    from rhseg_b200 import native
    lib = native.lib()
    ctx = ctypes.c_void_p()
    handle = (ctypes.c_ubyte * native.XCHG_HANDLE_BYTES)()
    assert lib.rhseg_xchg_create(0, 1, ctypes.byref(ctx), handle) < 0
    assert lib.rhseg_xchg_create(1024, 17, ctypes.byref(ctx), handle) != 0       # world too large
    assert lib.rhseg_xchg_create(1024, 1, ctypes.byref(ctx), handle) == 0
    out = torch.zeros(8, dtype=torch.float64, device="cuda")
    # not connected yet
    assert lib.rhseg_xchg_all_reduce(ctx, out.data_ptr(), 8, None, None, 0, out.data_ptr(), None) < 0
    assert lib.rhseg_xchg_connect(ctx, 1, bytes(handle)) != 0              # rank outside the world
    assert lib.rhseg_xchg_connect(ctx, 0, bytes(handle)) == 0              # world of one: sum == input
    src = torch.arange(8, dtype=torch.float64, device="cuda")
    part = torch.arange(5, dtype=torch.float32, device="cuda")
    res = torch.zeros(13, dtype=torch.float64, device="cuda")
    ptrs = (ctypes.c_void_p * 1)(part.data_ptr())
    cnts = (ctypes.c_long * 1)(5)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        assert lib.rhseg_xchg_all_reduce(ctx, src.data_ptr(), 8, ptrs, cnts, 1, res.data_ptr(), st) == 0
    torch.cuda.synchronize()
    assert torch.equal(res, torch.cat([src, part.double()]))
    assert lib.rhseg_xchg_all_reduce(ctx, src.data_ptr(), 2000, None, None, 0, res.data_ptr(), st) < 0   # over capacity
    s = ctypes.c_int(-1)
    assert lib.rhseg_xchg_status(ctx, ctypes.byref(s)) == 0 and s.value == 0
    assert lib.rhseg_xchg_destroy(ctx) == 0
