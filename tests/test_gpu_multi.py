"""Two-GPU parity of the time-sharded chain over NCCL (tools/shard_parity.py): halo exchange + loop-state ring must
reproduce the single-chain run bit for bit.  Needs >= 2 CUDA devices (gpurun --gpus 2); skipped otherwise."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ngpus():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.skipif(ngpus() < 2, reason="needs two CUDA devices")
def test_two_time_shards_equal_one_chain():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tools", "shard_parity.py"), "--steps", "2"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "shard parity ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.skipif(ngpus() < 2, reason="needs two CUDA devices")
@pytest.mark.parametrize("handoff", ["peer", "nccl"])
def test_bench_sharded_run_checks_itself_against_one_chain(handoff):
    """bench.py --gpus 2: ONE stream time-sharded over two ranks (ShardedChain: zero-copy state rings, correlator on its
    own chain), every sync hit gathered to rank 0 and compared with a single chain there; the run asserts the equality.
    Both forms of the loop-state hand-off: peer memory + stream memory operations, and NCCL send/recv."""
    import json
    env = dict(os.environ, GRB_SHARD_HANDOFF=handoff)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29542" if handoff == "peer" else "29543", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "3", "--warmup", "3",
           "--rows", "2500", "--sustain-seconds", "0", "--no-cpu", "--e2e-steps", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=400, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["n_gpus"] == 2 and line["parity"]["identical_to_single_chain"] is True and line["parity"]["sync_hits"] > 1000
    if handoff == "nccl":
        assert line["config"]["loop_state_handoff"] == "nccl send/recv"
    assert line["e2e"]["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] > 0
