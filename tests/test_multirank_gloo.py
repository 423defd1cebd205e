"""N>1 path on CPU: world_size 2 over gloo, tasks sharded by rank, one all-reduce of the meta-gradient.
(The kernels run on the host emulation here; the same worker runs with NCCL on GPUs: -m gpu below.)"""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _run(device, world, port):
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "multirank_worker.py"), device], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=900)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o[-3000:]}"
        assert f"rank {r} ok" in o


def test_two_ranks_gloo_sharded_tasks_match_single_process_gradient():
    _run("cpu", 2, 29611)


@pytest.mark.gpu
def test_two_ranks_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run("cuda", 2, 29612)
