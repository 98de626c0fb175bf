"""Worker of tests/test_gpu_parity.py::test_two_ranks_nccl -- launched with torchrun, one rank per GPU.

Every rank owns a contiguous, length-balanced range of the corpus (p normalised over the WHOLE corpus), attaches
the library's NCCL communicator (wfsa_dev_comm_init) and evaluates; structure counts, log-likelihood and gradient
are all-reduced inside the library as exact 64-bit integers, so every rank must hold the same bits, and they must
equal what ONE device computes on the whole corpus."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "w-fsa_b200", "python"))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import wfsa_b200 as W
    from oracle import oracle as O
    from wfsa_b200 import synth

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    kernel = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0      # e.g. 4 << 16: four pool slots, so that some strings overflow to the secondary kernel
    model = synth.make_model(256, 64, 8, 4, seed=11)
    low = model.lowered()
    offs, toks, w = model.corpus(6000, 32, 128, seed=12)
    toks = toks.copy()
    toks[offs[5]] = -1                                   # one unrecognised string
    low.set_tokens(offs, toks, w / w.sum())
    cuts = synth.balanced_ranges(offs, world)
    a, b = int(cuts[rank]), int(cuts[rank + 1])

    uid = np.zeros(W.UNIQUE_ID_BYTES, dtype=np.uint8)
    if rank == 0:
        assert W.lib().wfsa_dev_comm_unique_id(uid.ctypes.data_as(W.C.c_void_p)) == 0
    t = torch.from_numpy(uid).cuda()
    dist.broadcast(t, 0)
    dev = W.Device(low, device=local, force_kernel=kernel, accum_variant=variant, first=a, count=b - a)
    dev.comm_init(t.cpu().numpy().tobytes(), rank, world)
    rec, pc, used = dev.structure()                      # `used` is combined over all ranks
    params = np.concatenate([low.trans_param, low.emis_param])
    trimmed, n, _ = O.trim(low, np.array([r < 0 or used[r] for r in params]))
    dev.set_param_map(trimmed, n, rec)
    # several evaluations in a row (the exchange alternates between two packet buffers by epoch), through the host-buffer
    # call (one graph launch on the single-launch path) and through upload / launch / fetch
    for k in range(4):
        xk = np.random.RandomState(100 + k).normal(-1.0, 0.4, size=n)
        llk, _, gk = dev.eval(xk, want_logq=False)
        dev.upload_x(xk)
        dev.eval_launch()
        llr, gr = dev.eval_fetch()
        assert llk == llr and np.array_equal(gk, gr), "host-buffer call and resident evaluation differ"
        mk = torch.from_numpy(np.concatenate([[llk], gk])).cuda()
        rk = mk.clone()
        dist.broadcast(rk, 0)
        assert torch.equal(mk.view(torch.int64), rk.view(torch.int64)), "ranks disagree (evaluation %d)" % k
    x = np.random.RandomState(3).normal(-1.2, 0.6, size=n)
    ll, logq, grad = dev.eval(x)
    path = dev.info()["eval_path"]
    dev.close()

    # every rank holds the same bits
    mine = torch.from_numpy(np.concatenate([[ll], grad])).cuda()
    ref = mine.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(mine.view(torch.int64), ref.view(torch.int64)), "ranks disagree"
    if rank == 0:
        one = W.Device(low, device=local, force_kernel=kernel, accum_variant=variant)
        one_kernel = one.info()["kernel"]
        rec1, pc1, used1 = one.structure()
        assert np.array_equal(used1, used) and np.array_equal(rec1[a:b], rec)
        one.set_param_map(trimmed, n, rec1)
        ll1, logq1, grad1 = one.eval(x)
        one.close()
        if one_kernel in (1, 2, 3, 4):
            # per-string kernels round every posterior of every string: sums over shards are bitwise additive
            assert ll1 == ll and np.array_equal(grad1, grad), "sharded evaluation differs from the single-device one"
        else:
            # the segmented path rounds W_type * posterior, and W_type is summed per shard: equal up to the
            # fixed-point quantum (2^-54 per addend), not bitwise
            assert abs(ll1 - ll) <= 1e-13 * abs(ll) and np.allclose(grad1, grad, rtol=1e-12, atol=1e-15)
        assert np.allclose(logq1[a:b], logq, rtol=1e-14, atol=0, equal_nan=True)
        ltw, lew = low.edge_logweights(x, trimmed)
        _, olq, oee = O.dp_eval(low, ltw, lew)
        r = rec1.astype(bool)
        assert abs(ll - float(np.sum(low.p[r] * olq[r]))) <= 1e-10 * abs(ll)
        print("nccl_worker ok: world=%d kernel=%d eval_path=%d loglik=%.15g" % (world, kernel, path, ll), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
