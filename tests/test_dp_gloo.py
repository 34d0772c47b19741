"""CPU, world_size 2 (gloo): the data-parallel host logic — batch sharding, the flat gradient bucket layout, ONE
all-reduce(SUM) and the 1/W fold — reproduces the single-process full-batch gradients and Adam update.  The per-rank
"compute" is the numpy oracle standing in for the CUDA kernels (this test is about the plumbing around them)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from motionmixerconv_b200 import parallel as P_
        from oracle import mixer_np as O
        from tests.golden_util import Golden
        from tests.synthetic import synthetic_pose_windows
        g = Golden(case)
        c = g.cfg
        x, gt = synthetic_pose_windows(8, c["seq_len"], c["pred_len"], c["input_size"], scale="h36m", seed=21)
        xs, gs = P_.shard(torch.from_numpy(x), rank, world).numpy(), P_.shard(torch.from_numpy(gt), rank, world).numpy()
        orc = O.MlpMixerOracle(c, g.params)
        loss, dpred = O.mpjpe(orc.forward(xs), gs)
        grads, _ = orc.backward(dpred)
        keys = O.trainable_keys(g.params)
        offs, n = P_.flat_offsets([g.params[k].size for k in keys])
        bucket = torch.zeros(n)
        for k, o in zip(keys, offs):
            bucket[o:o + grads[k].size] = torch.from_numpy(np.ascontiguousarray(grads[k]).reshape(-1))
        scale = P_.allreduce_bucket(bucket)            # one collective for all gradients
        assert scale == 1.0 / world
        lt = torch.tensor([float(loss)])
        dist.all_reduce(lt)
        if rank == 0:
            np.savez(os.path.join(out_dir, "dp.npz"), bucket=(bucket * scale).numpy(), loss=lt.item() / world)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_bucket_allreduce_matches_full_batch(tmp_path):
    from motionmixerconv_b200 import parallel as P_
    from oracle import mixer_np as O
    from tests.golden_util import Golden
    from tests.synthetic import synthetic_pose_windows
    case, world = "mlp_odd_nose", 2
    mp.spawn(_worker, args=(world, _free_port(), case, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "dp.npz")
    g = Golden(case)
    c = g.cfg
    x, gt = synthetic_pose_windows(8, c["seq_len"], c["pred_len"], c["input_size"], scale="h36m", seed=21)
    orc = O.MlpMixerOracle(c, g.params, dtype=np.float64)
    loss, dpred = O.mpjpe(orc.forward(x), gt.astype(np.float64))
    grads, _ = orc.backward(dpred)
    keys = O.trainable_keys(g.params)
    offs, n = P_.flat_offsets([g.params[k].size for k in keys])
    assert got["bucket"].shape == (n,)
    scale = max(float(np.abs(v).max()) for v in grads.values())
    for k, o in zip(keys, offs):
        np.testing.assert_allclose(got["bucket"][o:o + grads[k].size], grads[k].reshape(-1), rtol=0, atol=2e-5 * scale, err_msg=k)
    assert abs(float(got["loss"]) - float(loss)) <= 1e-5 * abs(float(loss))


def test_shard_bounds_and_layout():
    from motionmixerconv_b200 import parallel as P_
    assert P_.shard_bounds(4096, 3, 8) == (1536, 2048)
    with pytest.raises(ValueError):
        P_.shard_bounds(10, 0, 4)
    offs, n = P_.flat_offsets([5, 8, 1])
    assert offs == [0, 8, 16] and n == 20 and all(o % P_.ALIGN == 0 for o in offs)
