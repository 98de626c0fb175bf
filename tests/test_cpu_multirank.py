"""world_size-2 gloo test (CPU) of the multi-GPU plan: the corpus is cut into
length-balanced contiguous ranges, every rank evaluates its shard with p normalised over the WHOLE
corpus, and [loglik, per-edge expected counts] are combined with ONE all-reduce of 64-bit
fixed-point integers -- exact, so any rank count gives bitwise identical results.  The per-rank
evaluation here is the CPU oracle (the device needs a GPU); the sharding, the quantisation rule
(round-to-nearest of value * 2^k, the same as kernels.cuh) and the integer all-reduce are what is
under test."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_strings, q):
    sys.path.insert(0, os.path.join(ROOT, "w-fsa_b200", "python"))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from wfsa_b200 import synth
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = synth.make_model(64, 16, 4, 3, seed=5)
    low = model.lowered()
    offs, toks, w = model.corpus(n_strings, 8, 40, seed=6)
    p = w / w.sum()
    low.set_tokens(offs, toks, p)
    cuts = synth.balanced_ranges(offs, world)
    a, b = cuts[rank], cuts[rank + 1]
    x = np.random.RandomState(1).normal(-1, 0.4, size=low.n_raw)
    ltw, lew = low.edge_logweights(x, np.arange(low.n_raw))
    _, lq, ee = O.dp_eval(low, ltw, lew, first=a, count=b - a, nthreads=1)
    k = 62 - 6                                           # max steps < 64
    fx = np.rint(ee * 2.0 ** k).astype(np.int64)         # NB: the device quantises every arc posterior, this
    ll = np.rint(np.sum(p[a:b] * lq) * 2.0 ** 44).astype(np.int64)   # test quantises per-rank sums: same all-reduce
    buf = torch.from_numpy(np.concatenate([[ll], fx]))
    dist.all_reduce(buf)
    tok_cnt = torch.tensor([int(offs[b] - offs[a])])
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, tok_cnt)
    if rank == 0:
        q.put((buf.numpy().copy(), [int(s) for s in sizes], cuts))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_integer_allreduce_is_exact(world):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "w-fsa_b200", "python"))
    from wfsa_b200 import synth
    from oracle import oracle as O
    n_strings = 600
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_strings, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    got, sizes, cuts = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    # ranges are contiguous, cover everything, and balanced by token count
    assert cuts[0] == 0 and cuts[-1] == n_strings and all(b >= a for a, b in zip(cuts, cuts[1:]))
    assert max(sizes) - min(sizes) <= 45                 # at most ~ one string apart
    # single-process reference: sum of the per-shard quantised values is what every rank must hold
    model = synth.make_model(64, 16, 4, 3, seed=5)
    low = model.lowered()
    offs, toks, w = model.corpus(n_strings, 8, 40, seed=6)
    p = w / w.sum()
    low.set_tokens(offs, toks, p)
    x = np.random.RandomState(1).normal(-1, 0.4, size=low.n_raw)
    ltw, lew = low.edge_logweights(x, np.arange(low.n_raw))
    tot = np.zeros(low.n_trans + low.n_emis + 1, dtype=np.int64)
    for a, b in zip(cuts, cuts[1:]):
        _, lq, ee = O.dp_eval(low, ltw, lew, first=a, count=b - a, nthreads=1)
        tot[1:] += np.rint(ee * 2.0 ** 56).astype(np.int64)
        tot[0] += np.rint(np.sum(p[a:b] * lq) * 2.0 ** 44).astype(np.int64)
    assert np.array_equal(got, tot)                       # bit for bit, whatever the reduction order
    _, lq, ee = O.dp_eval(low, ltw, lew, nthreads=1)
    assert np.allclose(got[1:] * 2.0 ** -56, ee, rtol=1e-12, atol=1e-15)
    assert abs(got[0] * 2.0 ** -44 - np.sum(p * lq)) < 1e-11


def test_balanced_ranges_edge_cases():
    sys.path.insert(0, os.path.join(ROOT, "w-fsa_b200", "python"))
    from wfsa_b200 import synth
    offs = np.array([0, 5, 5, 9, 100, 101], dtype=np.int64)
    for parts in (1, 2, 4, 8):
        cuts = synth.balanced_ranges(offs, parts)
        assert len(cuts) == parts + 1 and cuts[0] == 0 and cuts[-1] == 5
        assert all(b >= a for a, b in zip(cuts, cuts[1:]))
    assert synth.balanced_ranges(np.array([0], dtype=np.int64), 4) == [0, 0, 0, 0, 0]
