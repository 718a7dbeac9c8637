"""CPU, gloo, world_size 2: the N>1 host logic -- shard the loci, compute per rank, gather fixed-size call records to rank 0 --
reproduces the single-process result.  The per-rank compute stand-in here is the CPU oracle (tests may use it); on GPUs the same
sharding feeds sx_site_gl_germline and the gather is sx_gather_records over NCCL (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_sites, out_path):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import reflib
    import specgen
    from strelka_b200 import _abi as A
    from strelka_b200 import batch as B
    from strelka_b200.shard import shard_range

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(99)  # every rank sees the same full input and takes its shard
    pb = specgen.random_pileups(rng, n_sites, depth=20.0)
    a, b = shard_range(n_sites, rank, world)
    sub = B.PileupBatch(pb.site_off[a:b + 1] - pb.site_off[a], pb.calls[pb.site_off[a]:pb.site_off[b]], pb.ref_base[a:b])
    rec = reflib.ox_germline(A.default_params(), sub, True)
    local = torch.from_numpy(rec.view(np.uint8).reshape(b - a, -1).copy())
    # equal-count shards pad to the largest block so that one fixed-size gather suffices
    cap = (n_sites + world - 1) // world
    buf = torch.zeros((cap, local.shape[1]), dtype=torch.uint8)
    buf[: b - a] = local
    gathered = [torch.zeros_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, gathered, dst=0)
    if rank == 0:
        parts = []
        for r in range(world):
            ra, rb = shard_range(n_sites, r, world)
            parts.append(gathered[r][: rb - ra].numpy())
        allrec = np.concatenate(parts).reshape(-1).view(A.DIGT_RESULT_DT)
        np.save(out_path, allrec)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_compute_and_gather_matches_single_process(tmp_path):
    import torch.multiprocessing as mp

    sys.path.insert(0, HERE)
    import reflib
    import specgen
    from strelka_b200 import _abi as A

    n_sites, world = 501, 2
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(world, _free_port(), n_sites, out), nprocs=world, join=True)
    got = np.load(out)
    rng = np.random.default_rng(99)
    pb = specgen.random_pileups(rng, n_sites, depth=20.0)
    want = reflib.ox_germline(A.default_params(), pb, True)
    assert got.tobytes() == want.tobytes()


def test_shard_ranges_cover_and_balance():
    from strelka_b200.shard import gathered_offsets, shard_ranges

    for n in (0, 1, 7, 1000, 1_000_003):
        for w in (1, 2, 4, 8):
            rs = shard_ranges(n, w)
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1
            assert gathered_offsets(sizes)[-1] + sizes[-1] == n


def _enum_regions(seed, n_regions):
    import specgen

    rng = np.random.default_rng(seed)
    return [specgen.random_enum_region(rng, n_reads=int(rng.integers(1, 6)), cluster=bool(i % 2), n_keys=(1, 6)) for i in range(n_regions)]


def _enum_worker(rank, world, port, n_regions, out_path):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import reflib
    from strelka_b200 import batch as B
    from strelka_b200.shard import shard_range

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    regions = _enum_regions(123, n_regions)  # every rank sees the same loci and takes its shard: regions are independent
    a, b = shard_range(n_regions, rank, world)
    eb = B.EnumBatch(regions[a:b])
    out = reflib.ox_enumerate_alignments(eb, cap_alns=eb.n_reads * 64 + 64)  # the per-rank stand-in for sx_enumerate_alignments
    # what a rank contributes to the job-wide tally: reads, alignments, segments, keys, a checksum of the alignment positions
    local = torch.tensor([eb.n_reads, int(out.totals[0]), int(out.totals[1]), int(out.totals[2]), int(out.aln_pos[: int(out.totals[0])].astype(np.int64).sum())],
                         dtype=torch.int64)
    gathered = [torch.zeros_like(local) for _ in range(world)] if rank == 0 else None
    dist.gather(local, gathered, dst=0)
    if rank == 0:
        np.save(out_path, torch.stack(gathered).numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_enumeration_shards_by_region(tmp_path):
    """K7 (and with it the whole realignment chain) shards like everything else: regions are independent, each rank enumerates its
    contiguous block, no data-path collective; the gathered per-rank tallies add up to the single-process result."""
    import torch.multiprocessing as mp

    sys.path.insert(0, HERE)
    import reflib
    from strelka_b200 import batch as B

    n_regions, world = 41, 2
    out = str(tmp_path / "enum_tally.npy")
    mp.spawn(_enum_worker, args=(world, _free_port(), n_regions, out), nprocs=world, join=True)
    got = np.load(out).sum(axis=0)
    eb = B.EnumBatch(_enum_regions(123, n_regions))
    whole = reflib.ox_enumerate_alignments(eb, cap_alns=eb.n_reads * 64 + 64)
    want = [eb.n_reads, int(whole.totals[0]), int(whole.totals[1]), int(whole.totals[2]), int(whole.aln_pos[: int(whole.totals[0])].astype(np.int64).sum())]
    assert list(got) == want and want[1] > 300


def _variant_worker(rank, world, port, n_sites, root_capacity, out_path):
    """The N>1 step of the whole-path bench on the CPU: each rank genotypes its shard of positions, compacts the non-reference sites into
    sx_site_call records, and the records go to rank 0 by the protocol of sx_gatherv_records (csrc/sx_comm.cu): byte counts first, blocks at
    their prefix sums, and a too-small root buffer reported on EVERY rank instead of a hang."""
    import torch
    import torch.distributed as dist

    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import reflib
    import specgen
    from strelka_b200 import _abi as A
    from strelka_b200 import batch as B
    from strelka_b200.shard import gathered_offsets, shard_range

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(7)
    pb = specgen.random_pileups(rng, n_sites, depth=20.0)
    a, b = shard_range(n_sites, rank, world)
    sub = B.PileupBatch(pb.site_off[a:b + 1] - pb.site_off[a], pb.calls[pb.site_off[a]:pb.site_off[b]], pb.ref_base[a:b])
    rec = reflib.ox_germline(A.default_params(), sub, True)
    local = _site_calls(rec, np.diff(sub.site_off), first_pos=a)
    mine = torch.tensor([local.nbytes], dtype=torch.int64)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, mine)
    counts = [int(c) for c in counts]
    offs = gathered_offsets(counts)
    verdict = torch.tensor([1 if (rank != 0 or offs[-1] + counts[-1] <= root_capacity) else 0], dtype=torch.int64)
    dist.broadcast(verdict, 0)
    if not int(verdict):
        np.save(out_path + f".refused{rank}.npy", np.zeros(1))
    else:
        cap = max(counts)
        buf = torch.zeros(cap, dtype=torch.uint8)
        buf[: local.nbytes] = torch.from_numpy(local.view(np.uint8).reshape(-1).copy())
        got = [torch.zeros(cap, dtype=torch.uint8) for _ in range(world)] if rank == 0 else None
        dist.gather(buf, got, dst=0)
        if rank == 0:
            allb = np.zeros(offs[-1] + counts[-1], np.uint8)
            for r in range(world):
                allb[offs[r]: offs[r] + counts[r]] = got[r][: counts[r]].numpy()
            np.save(out_path, allb.view(A.SITE_CALL_DT))
    dist.barrier()
    dist.destroy_process_group()


def _site_calls(rec, n_calls, first_pos):
    """sx_site_call records of the positions whose most likely genotype is not the reference's (what sxp_variant_write_kernel compacts)"""
    from strelka_b200 import _abi as A

    keep = [i for i in range(len(rec)) if rec["is_computed"][i] and rec["genome"]["max_gt"][i] != rec["ref_gt"][i]]
    out = np.zeros(len(keep), A.SITE_CALL_DT)
    for k, i in enumerate(keep):
        out["pos"][k] = first_pos + i
        out["n_calls"][k] = n_calls[i]
        out["gl"][k] = rec[i]
    return out


def test_variant_records_gather_with_variable_block_sizes(tmp_path):
    import torch.multiprocessing as mp

    sys.path.insert(0, HERE)
    import reflib
    import specgen
    from strelka_b200 import _abi as A

    n_sites, world = 777, 2
    rng = np.random.default_rng(7)
    pb = specgen.random_pileups(rng, n_sites, depth=20.0)
    want = _site_calls(reflib.ox_germline(A.default_params(), pb, True), np.diff(pb.site_off), first_pos=0)
    assert 0 < len(want) < n_sites
    out = str(tmp_path / "variants.npy")
    mp.spawn(_variant_worker, args=(world, _free_port(), n_sites, want.nbytes, out), nprocs=world, join=True)
    assert np.load(out).tobytes() == want.tobytes()  # rank order == position order: no re-sort
    # a root buffer one record short: every rank is told, nobody waits in a collective
    out2 = str(tmp_path / "short.npy")
    mp.spawn(_variant_worker, args=(world, _free_port(), n_sites, want.nbytes - 1, out2), nprocs=world, join=True)
    assert not os.path.exists(out2) and all(os.path.exists(out2 + f".refused{r}.npy") for r in range(world))
