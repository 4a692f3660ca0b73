"""CPU suite: the N > 1 host logic (batch sharding, max-over-ranks timing) with world_size = 2 over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffpiso_b200 import sharding


def test_shard_bounds_cover_batch_exactly():
    for total in (1, 2, 7, 64, 65, 256):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                s, c = sharding.shard_bounds(total, world, r)
                seen += list(range(s, s + c))
            assert seen == list(range(total))
    with pytest.raises(ValueError):
        sharding.shard_bounds(8, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    batch = torch.arange(13 * 3, dtype=torch.float32).reshape(13, 3)
    mine = sharding.local_batch(batch)
    counts = sharding.gather_counts(mine.shape[0])
    times = sharding.max_over_ranks([10.0 + rank, 5.0 - rank])
    first = float(mine[0, 0])
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, counts, times, first, mine.shape[0]))


def test_two_rank_gloo_sharding_and_timing():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [[7, 6], [7, 6]]            # every rank sees the same whole-job counts
    assert [r[2] for r in res] == [[11.0, 5.0], [11.0, 5.0]]  # max over ranks
    assert [r[3] for r in res] == [0.0, 21.0] and [r[4] for r in res] == [7, 6]


def _train_worker(rank, world, port, q):
    """Two ranks, each with its own shard of samples: after allreduce_gradients + Adam both hold identical closure
    weights, equal to a single-process run over the whole batch."""
    from diffpiso_b200 import networks as N, training as T
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gen = torch.Generator().manual_seed(3)
    net, w, _ = N.initialise_fullyconv_network(None, "SAME", generator=gen)
    x = torch.randn(4, 12, 14, 4, generator=torch.Generator().manual_seed(5))
    mine = sharding.local_batch(x)
    opt = torch.optim.Adam(w, lr=1e-3)
    loss = T.training_iteration(opt, w, lambda: (net(mine) ** 2).sum() / x.shape[0] * world)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, float(loss), [t.detach().numpy().copy() for t in w]))


def test_two_rank_gloo_closure_gradient_allreduce():
    from diffpiso_b200 import networks as N, training as T
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=180) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for a, b in zip(res[0][2], res[1][2]):
        assert (a == b).all()
    # single process over the whole batch (mean over ranks of per-rank losses scaled by world == whole-batch loss)
    gen = torch.Generator().manual_seed(3)
    net, w, _ = N.initialise_fullyconv_network(None, "SAME", generator=gen)
    x = torch.randn(4, 12, 14, 4, generator=torch.Generator().manual_seed(5))
    opt = torch.optim.Adam(w, lr=1e-3)
    T.training_iteration(opt, w, lambda: (net(x) ** 2).sum() / x.shape[0])
    for a, b in zip(res[0][2], w):
        assert torch.allclose(torch.from_numpy(a), b.detach(), rtol=1e-4, atol=1e-6)


def test_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the oracle port on the host cores) prints one JSON line with the contract keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "piso_cell_updates_per_s_fwd_adjoint"
    assert line["unit"] == "cell-updates/s" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["config"]["workload"].startswith("decaying_turbulence_periodic_128x128")


def test_sample_groups_argument_checks_without_a_gpu():
    """`SampleGroups` (sample groups on concurrent CUDA streams) validates its inputs before it touches CUDA and has no
    CPU path; the group boundaries are the same contiguous split the ranks use."""
    import torch
    from diffpiso_b200 import sharding
    fn = lambda a: (a,)
    with pytest.raises(ValueError, match="at least one"):
        sharding.SampleGroups(fn, (), groups=2)
    with pytest.raises(ValueError, match="leading dimension"):
        sharding.SampleGroups(fn, (torch.zeros(4, 3), torch.zeros(3, 3)), groups=2)
    with pytest.raises(ValueError, match="CUDA"):
        sharding.SampleGroups(fn, (torch.zeros(4, 3),), groups=2, device="cpu")
    if not torch.cuda.is_available():
        with pytest.raises(ValueError, match="CUDA"):
            sharding.SampleGroups(fn, (torch.zeros(4, 3),), groups=2)
    assert [sharding.shard_bounds(10, 4, i) for i in range(4)] == [(0, 3), (3, 3), (6, 2), (8, 2)]
