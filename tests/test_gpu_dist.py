"""Data-parallel parity on real GPUs (needs >= 2): two ranks over NCCL reproduce the CPU oracle's
single-process result within 1e-3 (measured ~2e-6), eager and CUDA-graph modes."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


MODES = {'nccl': dict(TN_DP_HYBRID='0', TN_DP_FUSED='0'),       # one all-reduce, early bucket overlapped
         'nccl_eager': dict(TN_DP_HYBRID='0', TN_DP_FUSED='0', TN_GRAPH_NCCL='0'),   # graph, eager all-reduce, graph
         'hybrid': dict(TN_DP_HYBRID='1', TN_DP_FUSED='0'),     # default: two-shot peer-memory all-reduce of the early bucket + peer-memory tail
         'hybrid_nccl': dict(TN_DP_HYBRID='1', TN_DP_FUSED='0', TN_PEER_AR='0'),   # NCCL early bucket + peer-memory tail
         'fused': dict(TN_DP_HYBRID='0', TN_DP_FUSED='1')}      # everything over peer memory in the update kernel


@pytest.mark.gpu
@pytest.mark.parametrize('graph', [0, 1])
@pytest.mark.parametrize('mode', sorted(MODES))
def test_two_gpu_data_parallel_matches_oracle(graph, mode):
    """Two ranks reproduce the single-process oracle, for every implementation of the gradient
    exchange (see MODES), eager and captured."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port',
           str(29600 + graph + 2 * sorted(MODES).index(mode)),
           os.path.join(ROOT, 'tools', 'dp_check.py'), '--graph', str(graph)]
    env = dict(os.environ, **MODES[mode])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=180, cwd=ROOT, env=env)
    assert 'DP_CHECK_OK world=2' in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
