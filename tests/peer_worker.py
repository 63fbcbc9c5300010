"""Worker of tests/test_peer_exchange.py (run under torchrun, one process per GPU): the fused merge + exchange of the
sharded map search must give, on every rank, exactly the records of a single-GPU search over the whole map."""
import ctypes as C
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
orb = importlib.import_module("cooperative-orb-slam_b200")
synth = importlib.import_module("cooperative-orb-slam_b200.synth")


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    def gather_bytes(b):
        t = torch.tensor(list(b), dtype=torch.uint8, device=dev)
        allt = torch.empty((world, len(b)), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allt, t)
        return [allt[r].cpu().numpy().tobytes() for r in range(world)]

    peer = orb.PeerExchange(5000, rank, world, local, gather_bytes)
    L = orb.lib()
    cur = torch.cuda.current_stream(dev)
    checked = 0
    for nq, nm, seed in ((2000, 200000, 1), (1, 1000, 2), (333, 77777, 3), (5000, 40000, 4), (64, world - 1, 5)):
        m = synth.descriptors(max(nm, 1), seed=seed)[:nm]
        q, m, _ = synth.query_set(m, nq=nq, seed=seed + 10) if nm > nq else (synth.descriptors(nq, seed=seed + 20), m, None)
        if nm > 10:
            m[nm // 2] = m[3]; m[nm - 1] = m[3]; q[0] = m[3]      # ties across shards: the first index must win
        d_q = torch.from_numpy(q).to(dev); d_all = torch.from_numpy(np.ascontiguousarray(m).reshape(-1, 32)).to(dev)
        lo = nm * rank // world; hi = nm * (rank + 1) // world
        d_shard = d_all[lo:hi].contiguous() if hi > lo else torch.zeros((1, 32), dtype=torch.uint8, device=dev)
        full = torch.empty((nq, 4), dtype=torch.int32, device=dev)
        assert L.orbm_knn2_device(C.c_void_p(d_q.data_ptr()), nq, C.c_void_p(d_all.data_ptr()), nm, 0, C.c_void_p(full.data_ptr()), 0,
                                  C.c_void_p(cur.cuda_stream)) == 0
        for variant in (0, 3, 5):
            for rep in range(3):      # repeated calls exercise the two slot parities
                out = torch.full((nq, 4), -9, dtype=torch.int32, device=dev)
                peer.knn2(d_q.data_ptr(), nq, d_shard.data_ptr(), hi - lo, lo, out.data_ptr(), variant, cur.cuda_stream)
                torch.cuda.synchronize()
                assert peer.error() == 0
                assert bool((out == full).all().item()), (nq, nm, variant, rep, rank)
                checked += 1
    # ---- back-to-back searches alternating between two query sets and two ways of cutting the map (ranks finish at different
    # times): the pruning bounds that the ranks publish into each other's buffers (csrc/match.cu, the forwarder warp of the CTA
    # pair kernel) must never survive into the next search -- set A are near copies of map descriptors (bounds fall to a few
    # bits), set B random ones (a leaked A bound would prune away B's true neighbours).
    # Search A: queries near map descriptors over the whole 400k map, rank 0 holding only 1/8 of it (its shard is small enough to
    # publish bounds, the other ranks' shards are not: mixed decisions).  Search B: random queries over the first 150k descriptors,
    # cut evenly (every rank publishes and prunes with the shared bounds).
    nmA, nmB, nq = 400000, 150000, 2000
    m = synth.descriptors(nmA, seed=77)
    qa, m, _ = synth.query_set(m, nq=nq, seed=78)
    qb = synth.descriptors(nq, seed=79)
    d_all = torch.from_numpy(np.ascontiguousarray(m).reshape(-1, 32)).to(dev)
    cutsA = [0] + [nmA // 8 + (nmA - nmA // 8) * r // (world - 1) for r in range(world)]
    cutsB = [nmB * r // world for r in range(world + 1)]
    spans = [(cutsA[rank], cutsA[rank + 1]), (cutsB[rank], cutsB[rank + 1])]
    d_qs = [torch.from_numpy(x).to(dev) for x in (qa, qb)]
    fulls = [torch.empty((nq, 4), dtype=torch.int32, device=dev) for _ in range(2)]
    for d_q, full, n_map in zip(d_qs, fulls, (nmA, nmB)):
        assert L.orbm_knn2_device(C.c_void_p(d_q.data_ptr()), nq, C.c_void_p(d_all.data_ptr()), n_map, 0, C.c_void_p(full.data_ptr()), 0,
                                  C.c_void_p(cur.cuda_stream)) == 0

    def shard_ptr(k):
        return d_all[spans[k][0]:].data_ptr(), spans[k][1] - spans[k][0], spans[k][0]

    outs = [torch.full((nq, 4), -9, dtype=torch.int32, device=dev) for _ in range(2)]
    bad = torch.zeros(1, dtype=torch.int64, device=dev)
    for it in range(300):
        k = it & 1
        peer.knn2(d_qs[k].data_ptr(), nq, *shard_ptr(k), outs[k].data_ptr(), 5, cur.cuda_stream)
        bad += (outs[k] != fulls[k]).sum()          # queued on the same stream: no host synchronisation between the searches
    torch.cuda.synchronize()
    assert peer.error() == 0
    assert int(bad.item()) == 0, ("back-to-back alternating searches", rank, int(bad.item()))
    checked += 300
    # ---- two searches in flight: two streams, each with its own peer buffers -- the exchange of one search runs next to the
    # search kernel of the other (the small merge-exchange grid must make progress there) and the results stay those of one search
    peer2 = orb.PeerExchange(5000, rank, world, local, gather_bytes)
    side = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    for st in side:
        st.wait_stream(cur)
    bads = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(2)]
    for it in range(200):
        k = it & 1
        (peer, peer2)[k].knn2(d_qs[k].data_ptr(), nq, *shard_ptr(k), outs[k].data_ptr(), 5, side[k].cuda_stream)
        with torch.cuda.stream(side[k]):
            bads[k] += (outs[k] != fulls[k]).sum()
    torch.cuda.synchronize()
    assert peer.error() == 0 and peer2.error() == 0
    assert int(bads[0].item()) == 0 and int(bads[1].item()) == 0, ("two searches in flight", rank)
    checked += 200
    dist.barrier()
    peer2.close()
    peer.close()
    # ---- the key-frame message: every rank packs its own key frames, one fused exchange, every rank unpacks everybody's
    B, cap = 10, 1027
    flags = 1 | 2 | 4
    nbytes = orb.message_bytes(B, cap, flags)
    mpeer = orb.PeerExchange(nbytes // 16, rank, world, local, gather_bytes)

    def case(r):
        rng = np.random.default_rng(100 + r)
        k = np.zeros((B, cap), orb.KP_DTYPE)
        for f in ("x", "y", "size", "response"):
            k[f] = rng.uniform(0, 1500, (B, cap)).astype(np.float32)
        k["angle"] = rng.uniform(0, 360, (B, cap)).astype(np.float32); k["octave"] = rng.integers(0, 8, (B, cap)); k["class_id"] = -1
        ku = k.copy(); ku["x"] += np.float32(0.37)
        desc = rng.integers(0, 256, (B, cap, 32), dtype=np.uint8)
        cnt = rng.integers(0, cap + 1, B).astype(np.int32)
        ur = rng.uniform(-1, 700, (B, cap)).astype(np.float32); dep = rng.uniform(-1, 40, (B, cap)).astype(np.float32)
        mp = rng.normal(0, 5, (B, cap, 4)).astype(np.float32)
        return k, ku, desc, cnt, ur, dep, mp

    t = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy()).to(dev)
    k, ku, desc, cnt, ur, dep, mp = case(rank)
    d = [t(a) for a in (k, ku, desc, cnt, ur, dep, mp)]
    msg = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    allmsg = torch.zeros((world, nbytes), dtype=torch.uint8, device=dev)
    for rep in range(3):
        orb.pack_keyframes_device(d[0].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), B, cap, msg.data_ptr(), d[1].data_ptr(), d[4].data_ptr(),
                                  d[5].data_ptr(), d[6].data_ptr(), stream=cur.cuda_stream)
        mpeer.exchange_messages(msg.data_ptr(), nbytes, allmsg.data_ptr(), nbytes, cur.cuda_stream)
        torch.cuda.synchronize()
        assert mpeer.error() == 0
        for r in range(world):
            ek, eku, edesc, ecnt, eur, edep, emp = case(r)
            o = [torch.zeros_like(x) for x in d]
            orb.unpack_keyframes_device(allmsg[r].data_ptr(), B, cap, flags, o[0].data_ptr(), o[2].data_ptr(), o[3].data_ptr(), o[1].data_ptr(),
                                        o[4].data_ptr(), o[5].data_ptr(), o[6].data_ptr(), stream=cur.cuda_stream)
            torch.cuda.synchronize()
            gc = o[3].cpu().numpy().view(np.int32)
            assert np.array_equal(gc, ecnt), (rank, r)
            gk = o[0].cpu().numpy().view(orb.KP_DTYPE).reshape(B, cap); gd = o[2].cpu().numpy().reshape(B, cap, 32)
            gmp = o[6].cpu().numpy().view(np.float32).reshape(B, cap, 4)
            for b in range(B):
                n = ecnt[b]
                assert gk[b, :n].tobytes() == orb.quantize_lcm(ek[b, :n]).tobytes() and np.array_equal(gd[b, :n], edesc[b, :n])
                assert np.array_equal(gmp[b, :n], emp[b, :n])
    dist.barrier()
    mpeer.close()
    dist.destroy_process_group()
    if rank == 0:
        print("peer exchange ok: %d searches identical to the single-GPU result on %d ranks; key-frame messages of %d bytes exchanged and unpacked"
              % (checked, world, nbytes))


if __name__ == "__main__":
    main()
