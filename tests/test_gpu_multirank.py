"""-m gpu, needs >= 2 GPUs (skipped otherwise; run with `gpurun --gpus 2`): the frame-sharded multi-GPU path on real devices --
one rank per GPU, per-frame packed records written on the device, one NCCL all-gather, the gathered bytes verified on rank 0
(tools/multirank_check.py).  The host-side packing is covered without GPUs by tests/test_multiproc.py (gloo, world size 2)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_gathered_records_equal_each_ranks_frame():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = min(n, 4)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", os.path.join(ROOT, "tools", "multirank_check.py")], capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:])
    assert r.returncode == 0 and "MULTIRANK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
