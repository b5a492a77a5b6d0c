"""world_size-2 gloo test of the batch-sharding host logic (the N>1 path of bench.py / dist.py):
a crop's result must not depend on the rank it lands on, and the gathered [N,J,3] must equal the
single-process result.  The per-shard function here is the CPU oracle (the GPU path is covered by
tests/test_net_gpu.py::test_host_buffer_call_and_batch_invariance)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from metro_pose3d_b200.dist import shard_bounds


def test_shard_bounds_cover_batch_without_overlap():
    for n in (0, 1, 7, 256, 513):
        for g in (1, 2, 3, 8):
            b = [shard_bounds(n, g, r) for r in range(g)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(g - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from metro_pose3d_b200.dist import ShardedPoseEstimator
    from metro_pose3d_b200.joints import export_permutation
    from metro_pose3d_b200.weights import synth_head
    from oracle.metro_oracle import decode_ref
    perm = export_permutation('h36m')
    x = torch.from_numpy(synth_head(n, 8, 17, seed=3))

    def infer(shard):
        return torch.from_numpy(decode_ref(shard.numpy(), 17, 32, perm).astype(np.float32))

    est = ShardedPoseEstimator(infer, 17)
    out = est(x)
    # the asynchronous variant (side-stream gather on a GPU; same exchange on CPU tensors) and re-used buffers
    again = est.gather_async(infer(x[est.local_slice(n)]), n)
    est.wait()
    assert torch.equal(again, out)
    if rank == 0:
        q.put(out.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('n', [8, 7])
def test_two_rank_gather_equals_single_process(n):
    from metro_pose3d_b200.joints import export_permutation
    from metro_pose3d_b200.weights import synth_head
    from oracle.metro_oracle import decode_ref
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = decode_ref(synth_head(n, 8, 17, seed=3), 17, 32, export_permutation('h36m')).astype(np.float32)
    assert np.array_equal(got, ref)


def _count_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('bench_module', os.path.join(root, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    # ranks measure different step times (4.3 ms vs 4.9 ms): a rank-local count would be 232 vs 204 steps
    ms, ext = bench.agree_on_time_and_steps(4.3 if rank == 0 else 4.9, 20, world, dist, torch.device('cpu'))
    q.put((rank, ms, ext))
    dist.barrier()
    dist.destroy_process_group()


def test_bench_ranks_agree_on_the_extension_step_count():
    """Every bench step ends in a collective: all ranks must loop the same number of times (regression test for the
    round-2 hang of the 8-GPU run, where each rank derived the count from its own timing)."""
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_count_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][1] == got[1][1] == 4.9 and got[0][2] == got[1][2] == 204
