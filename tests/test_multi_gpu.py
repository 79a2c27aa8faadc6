"""BASELINE config 4 on real GPUs (needs >= 2 B200 in the box: `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`;
skipped on a one-GPU box): the plane in spanwise slabs, one process per GPU, the library's own NCCL hand-off of the finished plane
to rank 0 -- driven (a) from plain C through include/dfb200.h, (b) from Python through parallel.SlabFilter."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "tests", "_bin")
LIBDIR = os.path.join(ROOT, "digital-filtering_b200", "lib")


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4, 8])
def test_gather_through_the_c_abi(dfb, world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    os.makedirs(BIN, exist_ok=True)
    exe = os.path.join(BIN, "comm_c_test")
    subprocess.run(["/usr/bin/gcc", "-std=c11", "-Wall", "-D_POSIX_C_SOURCE=200809L", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "comm_c_test.c"), "-L" + LIBDIR, "-ldfb200", "-Wl,-rpath," + LIBDIR, "-o", exe],
                   check=True, capture_output=True)
    r = subprocess.run([exe, str(world)], capture_output=True, text=True, timeout=180)
    assert r.returncode == 0, (r.stdout, r.stderr[-2000:])
    tok = [ln for ln in r.stdout.splitlines() if ln.startswith("OK ")][-1].split()      # (NCCL prints its version banner on stdout)
    assert tok[0] == "OK" and int(tok[1]) == world and int(tok[4]) > 0


@pytest.mark.parametrize("transport", ["p2p", "nccl"])
@pytest.mark.parametrize("world", [2])
def test_slab_filter_python_caller(dfb, world, transport, monkeypatch):
    """both transports of the hand-off: peer-to-peer copies (CUDA IPC + copy engines, the default) and ncclSend/ncclRecv + assembly"""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    monkeypatch.setenv("DFB_GATHER_TRANSPORT", transport)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "slab_worker.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "SLAB_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])
