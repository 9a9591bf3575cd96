"""CPU, world_size 2, gloo: the N>1 host logic of the data-parallel path (shard -> local work -> gather)."""
import os

import torch
import torch.multiprocessing as mp

from vllm_omni_b200.diffusion.distributed import parallel_state as ps


def _worker(rank, world, port, tp, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    ps.init_distributed_environment(world_size=world, rank=rank, backend="gloo")
    ps.initialize_model_parallel(data_parallel_size=world // tp, tensor_parallel_size=tp, backend="gloo")
    dp, dpr = ps.get_data_parallel_world_size(), ps.get_data_parallel_rank()
    total = 5
    counts = [ps.shard_range(total, r, dp)[1] - ps.shard_range(total, r, dp)[0] for r in range(dp)]
    lo, hi = ps.shard_range(total, dpr, dp)
    local = torch.arange(lo, hi, dtype=torch.float32).view(-1, 1, 1).expand(-1, 3, 4).contiguous() * 10
    out = ps.gather_to_rank0(local, counts) if ps.get_tensor_model_parallel_rank() == 0 or tp == 1 else None
    if tp > 1:
        t = torch.full((4,), float(rank))
        torch.distributed.all_reduce(t, group=ps.get_tp_group())
        q.put(("tp", rank, t.tolist()))
    if tp == 1 and out is not None and dpr == 0:
        q.put(("gather", rank, out[:, 0, 0].tolist()))
    ps.destroy_distributed_env()


def _run(world, tp, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, tp, q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in range(1 if tp == 1 else world)]
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    return res


def test_dp2_gather_gloo():
    res = _run(2, 1, 29631)
    assert res == [("gather", 0, [0.0, 10.0, 20.0, 30.0, 40.0])]


def test_tp2_allreduce_group_gloo():
    res = sorted(_run(2, 2, 29632))
    assert res[0][2] == [1.0] * 4 and res[1][2] == [1.0] * 4


def _cfg_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    ps.init_distributed_environment(world_size=world, rank=rank, backend="gloo")
    ps.initialize_model_parallel(data_parallel_size=1, tensor_parallel_size=1, cfg_parallel_size=2, backend="gloo")
    assert ps.get_cfg_parallel_world_size() == 2 and ps.get_cfg_parallel_rank() == rank and ps.get_data_parallel_rank() == 0
    # each rank holds the noise prediction of ITS branch; after the exchange both hold [positive, negative]
    pos, neg = ps.cfg_all_gather(torch.full((2, 3), float(10 + rank)))
    q.put(("cfg", rank, pos[0, 0].item(), neg[0, 0].item()))
    ps.destroy_distributed_env()


def test_cfg_parallel_exchange_gloo():
    """CFG-parallel host logic: group layout and the per-step all-gather of the two branches' predictions."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_cfg_worker, args=(r, 2, 29633, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert res == [("cfg", 0, 10.0, 11.0), ("cfg", 1, 10.0, 11.0)]


def test_rank_layout_dp_cfg_tp():
    """rank = (dp * cfg + c) * tp + t  (the reference's "tp-sp-pp-cfg-dp" order, parallel_state.py:659)."""
    st = ps._STATE
    saved = (st.rank, st.tp_size, st.cfg_size, st.dp_size)
    try:
        st.tp_size, st.cfg_size, st.dp_size = 2, 2, 2
        seen = set()
        for r in range(8):
            st.rank = r
            seen.add((ps.get_data_parallel_rank(), ps.get_cfg_parallel_rank(), ps.get_tensor_model_parallel_rank()))
        assert seen == {(d, c, t) for d in range(2) for c in range(2) for t in range(2)}
        st.rank = 5
        assert (ps.get_data_parallel_rank(), ps.get_cfg_parallel_rank(), ps.get_tensor_model_parallel_rank()) == (1, 0, 1)
    finally:
        st.rank, st.tp_size, st.cfg_size, st.dp_size = saved


def _sp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    ps.init_distributed_environment(world_size=world, rank=rank, backend="gloo")
    ps.initialize_model_parallel(data_parallel_size=1, tensor_parallel_size=1, ulysses_degree=2, backend="gloo")
    assert ps.get_sequence_parallel_world_size() == 2 and ps.get_sequence_parallel_rank() == rank
    assert ps.get_data_parallel_rank() == 0 and ps.get_cfg_parallel_rank() == 0
    t = torch.full((3,), float(rank + 1))
    torch.distributed.all_reduce(t, group=ps.get_sp_group())
    flag = ps.any_rank_in_model_group(rank == 1)   # the attention-guard decision is shared by the SP group
    q.put(("sp", rank, t.tolist(), flag))
    ps.destroy_distributed_env()


def test_sequence_parallel_group_gloo():
    """SP (ulysses_degree) group layout and the group-wide OR used by the attention guard."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sp_worker, args=(r, 2, 29634, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert res == [("sp", 0, [3.0] * 3, True), ("sp", 1, [3.0] * 3, True)]


def test_rank_layout_with_sp():
    """rank = ((dp * cfg + c) * sp + s) * tp + t."""
    st = ps._STATE
    saved = (st.rank, st.tp_size, st.cfg_size, st.dp_size, st.sp_size)
    try:
        st.tp_size, st.cfg_size, st.dp_size, st.sp_size = 1, 2, 2, 2
        seen = set()
        for r in range(8):
            st.rank = r
            seen.add((ps.get_data_parallel_rank(), ps.get_cfg_parallel_rank(), ps.get_sequence_parallel_rank()))
        assert seen == {(d, c, s_) for d in range(2) for c in range(2) for s_ in range(2)}
        st.rank = 6
        assert (ps.get_data_parallel_rank(), ps.get_cfg_parallel_rank(), ps.get_sequence_parallel_rank()) == (1, 1, 0)
    finally:
        st.rank, st.tp_size, st.cfg_size, st.dp_size, st.sp_size = saved
