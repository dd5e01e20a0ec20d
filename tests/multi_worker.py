"""Worker for the world_size > 1 tests (launched by torch.distributed.run).

mode "cpu" (gloo, no GPU): checks the HOST-SIDE multi-rank logic of the product library — layout
ownership, per-rank motion items and the aggregated per-peer halo messages (sgc_halo_describe) — by
moving real data with numpy pack/unpack + gloo send/recv and comparing the result with the oracle's
single-rank exchange of the same global field.

mode "gpu" (nccl): the product path: slab-partitioned device levels, sgc_heat_run with NCCL halo
exchange overlapped with interior boxes, bit-exact against the oracle's full-domain run.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from _loadpkg import load_package  # noqa: E402
import pyoracle  # noqa: E402


def region_view(fab, fablo, lo, hi, comp):
    """view of fab[comp, k, j, i] over the index region [lo, hi]"""
    s = [slice(lo[d] - fablo[d], hi[d] - fablo[d] + 1) for d in range(3)]
    return fab[comp, s[2], s[1], s[0]]


def run_cpu(rank, world):
    sgc = load_package()
    O = pyoracle.Oracle()
    cases = [((0, 0, 0), (15, 15, 31), (4, 4, 4), 1, 2, 7, 0), ((0, 0, 0), (7, 7, 15), (4, 4, 4), 1, 1, 0, 0),
             ((-2, 3, 0), (9, 10, 15), (6, 4, 4), 2, 3, 5, 8)]
    for dlo, dhi, mb, ng, nc, per, trim in cases:
        dbl = sgc.DisjointBoxLayout(sgc.Box(dlo, dhi), mb, world, rank)
        allboxes, owner, _ = O.layout(dlo, dhi, mb, world)
        assert np.array_equal(dbl.boxes, allboxes) and np.array_equal(dbl.owner, owner)
        # global field, identical on every rank
        rng = np.random.default_rng(123)
        glob = rng.standard_normal(O.level_size(allboxes, ng, nc))
        goff = O.level_offsets(allboxes, ng, nc)
        want = glob.copy()
        O.level_exchange(allboxes, ng, nc, O.copier_items(dlo, dhi, mb, ng, per, trim), want)
        # my slab
        b0, b1 = dbl.localIdxBegin(), dbl.localIdxEnd()
        fabs = {}
        for g in range(b0, b1):
            n = allboxes[g][3:] - allboxes[g][:3] + 1 + 2 * ng
            fabs[g] = glob[goff[g]:goff[g] + nc * int(np.prod(n))].reshape(nc, n[2], n[1], n[0]).copy()
        src = {g: f.copy() for g, f in fabs.items()}       # exchange reads pre-exchange valid data
        items = dbl.copier_items(ng, per, trim)
        for it in items[items["send_proc"] == rank]:         # local items
            d = region_view(fabs[int(it["recv_box"])], allboxes[it["recv_box"]][:3] - ng, it["recv_lo"], it["recv_hi"], slice(None))
            d[...] = region_view(src[int(it["send_box"])], allboxes[it["send_box"]][:3] - ng, it["src_lo"], it["src_hi"], slice(None))
        # remote items: one aggregated message per peer, regions in linearOut order comp->k->j->i
        reqs, recvbufs = [], []
        for peer, recv, send in dbl.halo_lists(ng, per, trim):
            out = [region_view(src[int(it["send_box"])], allboxes[it["send_box"]][:3] - ng, it["src_lo"], it["src_hi"], slice(None)).ravel()
                   for it in send]
            sbuf = torch.from_numpy(np.concatenate(out)) if out else torch.zeros(0, dtype=torch.float64)
            nrecv = int(sum(np.prod(it["recv_hi"] - it["recv_lo"] + 1) for it in recv)) * nc
            rbuf = torch.zeros(nrecv, dtype=torch.float64)
            reqs.append(dist.isend(sbuf, peer))
            reqs.append(dist.irecv(rbuf, peer))
            recvbufs.append((recv, rbuf))
        for r in reqs:
            r.wait()
        for recv, rbuf in recvbufs:
            buf, pos = rbuf.numpy(), 0
            for it in recv:
                d = region_view(fabs[int(it["recv_box"])], allboxes[it["recv_box"]][:3] - ng, it["recv_lo"], it["recv_hi"], slice(None))
                d[...] = buf[pos:pos + d.size].reshape(d.shape)
                pos += d.size
            assert pos == len(buf)
        for g in range(b0, b1):
            w = want[goff[g]:goff[g] + fabs[g].size]
            assert np.array_equal(fabs[g].ravel(), w), (rank, g, dlo, dhi)
    print("rank %d: cpu multi-rank host logic ok" % rank, flush=True)


def run_gpu(rank, world, local_rank):
    sgc = load_package()
    O = pyoracle.Oracle()
    sgc.init(local_rank)
    ids = [sgc.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    sgc.comm_init(world, rank, ids[0])
    cases = (((0, 0, 0), (31, 31, 32 * world - 1), (16, 16, 16), 7, 0.125, 12),
             ((0, 0, 0), (63, 63, 64 * world - 1), (64, 64, 64), 7, 1.0 / 6.0, 6),
             ((0, 0, 0), (127, 63, 128 * world - 1), (64, 64, 64), 7, 0.2, 11),
             ((0, 0, 0), (127, 63, 128 * world - 1), (64, 64, 64), 3, 0.2, 10),   # z ends: physical
             ((0, 0, 0), (63, 127, 64 * world - 1), (64, 64, 64), 0, 0.125, 9),    # no periodicity
             ((0, 0, 0), (63, 31, 64 * world - 1), (32, 32, 32), 3, 0.125, 9),
             ((0, 0, 0), (15, 15, 8 * world - 1), (8, 8, 8), 0, 0.125, 5),
             ((0, 0, 0), (11, 7, 10 * world - 1), (6, 4, 5), 3, 0.2, 4))
    # every case three times: z halo read out of the neighbour's memory (default), the NCCL schedule (SGC_P2P=0:
    # pack, one send/recv pair per peer, each peer's message unpacked as it arrives — LevelData.H:507-536), and
    # the all-peer-memory schedule (SGC_PULL=1)
    for p2p_env, (dlo, dhi, mb, per, f, nstep) in [(e, c) for e in (None, "0", "pull") for c in cases]:
        os.environ.pop("SGC_P2P", None)
        os.environ.pop("SGC_PULL", None)
        if p2p_env == "pull":          # the last steps' exchanges read the neighbours' arenas too (comm_p2p_pull)
            os.environ["SGC_PULL"] = "1"
        elif p2p_env is not None:
            os.environ["SGC_P2P"] = p2p_env
        dbl = sgc.DisjointBoxLayout(sgc.Box(dlo, dhi), mb, world, rank)
        allboxes, _, _ = O.layout(dlo, dhi, mb)
        u0 = O.fill_hash(allboxes, 1, 0.0)
        goff = O.level_offsets(allboxes, 1, 1)
        want = u0.copy()
        O.level_heat(dlo, dhi, mb, 1, per, 0, f, nstep, want, 1)
        b0, b1 = dbl.localIdxBegin(), dbl.localIdxEnd()
        end = goff[b1] if b1 < len(goff) else len(u0)
        mine = u0[goff[b0]:end].copy()
        a, b = sgc.LevelData(dbl, 1, 1), sgc.LevelData(dbl, 1, 1)
        a.upload(mine)
        b.upload(mine)
        cp = sgc.Copier().defineExchangeLD(a, per, 0)
        # single exchange first (sgc_exchange through NCCL), against the oracle's global exchange
        a.exchange(cp)
        w1 = u0.copy()
        O.level_exchange(allboxes, 1, 1, O.copier_items(dlo, dhi, mb, 1, per, 0), w1)
        assert np.array_equal(a.download(), w1[goff[b0]:end]), ("exchange", rank, mb)
        a.upload(mine)
        sgc.profile_enable(True)
        sgc.profile_read(2)
        res = sgc.heat_run(a, b, cp, f, nstep)
        n2, _ = sgc.profile_read(2)
        sgc.profile_enable(False)
        got = res.download()
        assert np.array_equal(got, want[goff[b0]:end]), ("heat_run", rank, mb)
        if mb in ((64, 64, 64), (32, 32, 32)) and os.environ.get("SGC_P2P", "1") != "0" and not os.environ.get("SGC_NO_T2"):
            # the two-step passes pull their z tiles out of the neighbouring GPU's memory
            P = (nstep - 2) // 2            # an odd count ends through the third level (mapped by the neighbours too)
            assert n2 == (P if (P % 2 == 0 or P >= 3) else P & ~1), ("two-step passes", n2, nstep)
        if world > 1:
            assert cp.numRemoteItem() > 0
    os.environ.pop("SGC_P2P", None)
    os.environ.pop("SGC_PULL", None)
    sgc.sync()
    dist.barrier()
    sgc.comm_finalize()
    print("rank %d: gpu multi-rank parity ok" % rank, flush=True)


if __name__ == "__main__":
    mode = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if mode == "cpu":
        dist.init_process_group(backend="gloo", rank=rank, world_size=world)
        run_cpu(rank, world)
    else:
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="cpu:gloo,cuda:nccl", rank=rank, world_size=world)
        run_gpu(rank, world, local_rank)
    dist.barrier()
    dist.destroy_process_group()
