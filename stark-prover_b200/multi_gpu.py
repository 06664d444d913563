"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed for plumbing.

What shards (and nothing else does):
  * independent trace columns -- column c goes to rank c mod G; each rank runs iNTT -> coset NTT -> Merkle
    tree per column with no data-path collective, then the 32-byte roots are all-gathered
    (`commit_columns`, BASELINE cfg4);
  * one large column -- contiguous leaf ranges of size N/G are exact subtrees of the rs_merkle tree shape
    (G a power of two): every rank hashes its range, subtree roots are gathered and the top log2(G)
    levels are finished identically on every rank (`combine_subtree_roots`, `commit_leaf_ranges`);
  * the four-step NTT for domains too large for one launch chain (`four_step_lde`, BASELINE cfg5): local
    batched transforms, one all-to-all transpose, local transforms, and a second exchange that puts the
    evaluations back into natural order in contiguous blocks so that leaf ranges can be hashed locally.

The collective calls work with any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests,
where the compute callbacks are injected).
"""
from __future__ import annotations

import hashlib
from typing import Callable, Optional, Sequence

import numpy as np


def shard_columns(n_cols: int, rank: int, world: int) -> list[int]:
    """Columns owned by `rank`: c mod world == rank."""
    return [c for c in range(n_cols) if c % world == rank]


def _dist():
    import torch.distributed as dist
    return dist


def _comm_device(group=None):
    import torch
    dist = _dist()
    backend = dist.get_backend(group) if dist.is_initialized() else "gloo"
    return torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")


def all_gather_bytes(local: np.ndarray, group=None) -> np.ndarray:
    """All-gathers a uint8 array of identical shape on every rank; returns [world, *shape]."""
    import torch
    dist = _dist()
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local[None, ...].copy()
    dev = _comm_device(group)
    t = torch.from_numpy(np.array(local, copy=True)).to(dev)
    out = torch.empty((dist.get_world_size(group),) + tuple(t.shape), dtype=t.dtype, device=dev)
    dist.all_gather_into_tensor(out.view(-1), t.view(-1), group=group)
    return out.cpu().numpy()


def commit_columns(n_cols: int, commit_fn, rank: int, world: int, group=None) -> list[bytes]:
    """Column-parallel commitment: `commit_fn(c)` returns the 32-byte Merkle root of column c (LDE + tree on
    this rank's GPU).  Returns the roots of ALL columns, in column order, identical on every rank.
    `commit_fn` may be a list of callables (one per context/stream on this GPU): the rank's columns are then
    pipelined over that many host threads, so one column's upload and tree tail overlap another's bulk hashing."""
    mine = shard_columns(n_cols, rank, world)
    per_rank = -(-n_cols // world)
    local = np.zeros((per_rank, 32), dtype=np.uint8)
    fns = list(commit_fn) if isinstance(commit_fn, (list, tuple)) else [commit_fn]
    if len(fns) == 1:
        roots = [fns[0](c) for c in mine]
    else:
        import threading
        roots = [None] * len(mine)

        def run(k):
            for slot in range(k, len(mine), len(fns)):
                roots[slot] = fns[k](mine[slot])
        ths = [threading.Thread(target=run, args=(k,)) for k in range(len(fns))]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
    for slot, root in enumerate(roots):
        assert len(root) == 32
        local[slot] = np.frombuffer(root, dtype=np.uint8)
    allr = all_gather_bytes(local, group)                       # [world, per_rank, 32]
    return [allr[c % world, c // world].tobytes() for c in range(n_cols)]


def gpu_column_committer(sp, ctx, get_column: Callable[[int], np.ndarray], log_blowup: int, offset_in: int, offset_out: int,
                         keep: Optional[dict] = None) -> Callable[[int], bytes]:
    """commit_fn for `commit_columns` on a GPU context: trace column (evaluations on offset_in*<g>) ->
    coset LDE on offset_out*<h> -> MerkleTree (reference: interpolate + evaluate + MerkleTree::new)."""
    def commit(c: int) -> bytes:
        col = ctx.upload(get_column(c))
        lde = ctx.coset_lde_dev(col, offset_in, log_blowup, offset_out)
        col.free()
        tree = sp.MerkleTree.new(ctx, lde)
        root = tree.root_bytes()
        if keep is not None:
            keep[c] = (lde, tree)
        else:
            tree.free(); lde.free()
        return root
    return commit


def combine_subtree_roots(roots: Sequence[bytes]) -> bytes:
    """Finishes a tree whose leaves were hashed in G equal contiguous ranges: the G subtree roots are the
    nodes of one level; pairs are hashed upward (rs_merkle rule; lone node promoted).  Host-side: G <= 8."""
    level = list(roots)
    while len(level) > 1:
        nxt = [hashlib.sha256(level[i] + level[i + 1]).digest() for i in range(0, len(level) - 1, 2)]
        if len(level) & 1:
            nxt.append(level[-1])
        level = nxt
    return level[0]


def top_path(roots: Sequence[bytes], owner: int) -> bytes:
    """Authentication-path bytes for the levels above the subtree roots (sibling digests bottom -> top)."""
    out, level, j = b"", list(roots), owner
    while len(level) > 1:
        sib = j ^ 1
        if sib < len(level):
            out += level[sib]
        nxt = [hashlib.sha256(level[i] + level[i + 1]).digest() for i in range(0, len(level) - 1, 2)]
        if len(level) & 1:
            nxt.append(level[-1])
        level, j = nxt, j >> 1
    return out


def commit_leaf_ranges(subtree_root_fn: Callable[[], bytes], rank: int, world: int, group=None) -> tuple[bytes, list[bytes]]:
    """One big column whose leaves are split into `world` equal contiguous ranges (world a power of two and
    the range a power of two, so each range is an exact subtree).  Returns (root, all subtree roots)."""
    assert world & (world - 1) == 0, "leaf-range sharding needs a power-of-two world size"
    local = np.frombuffer(subtree_root_fn(), dtype=np.uint8).reshape(1, 32)
    allr = all_gather_bytes(local, group)
    subs = [allr[r, 0].tobytes() for r in range(world)]
    return combine_subtree_roots(subs), subs


# ------------------------------------------------------------------------------------------------ four-step NTT
def four_step_plan(log_n: int, world: int) -> tuple[int, int]:
    """Splits 2^log_n = N1 * N2 with both factors divisible by the world size."""
    a = log_n // 2
    b = log_n - a
    g = world.bit_length() - 1
    assert world == 1 << g and a >= g and b >= g, "four-step NTT: world must be a power of two <= sqrt(N)"
    return a, b          # N1 = 2^a (high digit of the input index), N2 = 2^b


def four_step_scatter_input(coeffs: np.ndarray, log_n: int, rank: int, world: int) -> np.ndarray:
    """Host-side input layout for rank `rank`: columns n2 in its slice, each column contiguous over n1:
    A[n2'][n1] = x[n1*N2 + rank*N2/G + n2'] (x zero-padded to 2^log_n)."""
    a, b = four_step_plan(log_n, world)
    n1, n2 = 1 << a, 1 << b
    x = np.zeros(1 << log_n, dtype=np.uint64)
    x[: len(coeffs)] = coeffs
    m = x.reshape(n1, n2)
    w = n2 // world
    return np.ascontiguousarray(m[:, rank * w:(rank + 1) * w].T)       # [n2/G][n1]


def four_step_reference(coeffs: np.ndarray, log_n: int, offset: int, ntt_fn, modulus: int, omega: int, world: int) -> list[np.ndarray]:
    """Pure-numpy model of the distributed data flow (used by the CPU tests to pin the index algebra):
    returns, for each rank, its natural-order block of the evaluations of `coeffs` on offset*<omega>.
    `ntt_fn(a, log_m, w)` is a size-2^log_m forward NTT with root w (natural in, natural out)."""
    a, b = four_step_plan(log_n, world)
    n1, n2, n = 1 << a, 1 << b, 1 << log_n
    x = np.zeros(n, dtype=object)
    pw = 1
    for j, c in enumerate(coeffs):
        x[j] = int(c) * pw % modulus
        pw = pw * offset % modulus
    m = x.reshape(n1, n2)
    w1, w2 = pow(omega, n2, modulus), pow(omega, n1, modulus)
    # phase A: column transforms over n1, then the twiddle w^(n2*k1)
    y = np.zeros((n1, n2), dtype=object)
    for c in range(n2):
        col = ntt_fn(np.array([int(v) for v in m[:, c]], dtype=np.uint64), a, w1)
        for k1 in range(n1):
            y[k1, c] = int(col[k1]) * pow(omega, c * k1, modulus) % modulus
    # phase C: row transforms over n2: X[k1 + N1*k2]
    out = np.zeros(n, dtype=np.uint64)
    for k1 in range(n1):
        row = ntt_fn(np.array([int(v) for v in y[k1, :]], dtype=np.uint64), b, w2)
        out[k1 + n1 * np.arange(n2)] = row
    blk = n // world
    return [out[r * blk:(r + 1) * blk] for r in range(world)]


# ---- the same flow on the GPUs ------------------------------------------------------------------------
def _as_torch(vec):
    import torch
    return torch.as_tensor(vec, device=torch.device("cuda", torch.cuda.current_device()))


class FourStepLDE:
    """Distributed coset evaluation X[k] = sum_j c_j (offset*w^k)^j of size N = N1*N2 over `world` ranks.

    Index algebra (n = n1*N2 + n2 with n1 the high digit, k = k1 + N1*k2):
        X[k1 + N1*k2] = sum_{n2} w_{N2}^(n2 k2) * w_N^(n2 k1) * sum_{n1} x[n1*N2 + n2] w_{N1}^(n1 k1)
    phase A  rank r owns the columns n2 in its slice ([n2'][n1], each column contiguous): coset scaling,
             batched size-N1 NTTs, twiddle w_N^(n2*k1)                                  (local kernels)
    exch. 1  all-to-all: rank s receives the rows k1 in its slice                        (NCCL / NVLink)
    phase C  batched size-N2 NTTs over n2 for the local rows                             (local kernels)
    exch. 2  all-to-all back to natural order: rank t receives k in [t*N/G, (t+1)*N/G)   (NCCL / NVLink)
    Each exchange moves 4*N/G bytes per rank, (G-1)/G of it over NVLink.  The coefficients are taken from a
    device vector that every rank holds (they are 1/blowup of the domain); the local transposes are strided
    device copies straight into the vector the next kernel works on.
    """

    def __init__(self, sp, ctx, log_n: int, offset: int, rank: int, world: int):
        self.sp, self.ctx, self.log_n, self.offset, self.rank, self.world = sp, ctx, log_n, offset, rank, world
        self.a, self.b = four_step_plan(log_n, world)
        self.n1, self.n2 = 1 << self.a, 1 << self.b
        self.omega = ctx.root_of_unity(log_n)
        self._keep = []

    def _handoff_to_lib(self):
        import torch
        torch.cuda.current_stream().synchronize()

    def phase_a(self, coeffs):
        """coeffs: host array or device Vec of the (<= N) coefficients -> torch int32 tensor
        [world (dest)][n2/G][n1/G] ready for the first exchange."""
        ctx, w = self.ctx, self.n2 // self.world
        cv = coeffs if hasattr(coeffs, "device_ptr") else ctx.upload(coeffs)
        n_c = len(cv)
        rows = -(-n_c // self.n2)                              # rows of the [N1][N2] matrix that hold coefficients
        v = ctx.zeros(w * self.n1)                             # [n2'][n1]
        ctx.sync()
        full = rows if rows * self.n2 == n_c else rows - 1    # complete rows
        src = _as_torch(cv)
        dst = _as_torch(v).view(w, self.n1)
        if full:
            dst[:, :full] = src[: full * self.n2].view(full, self.n2)[:, self.rank * w:(self.rank + 1) * w].t()
        if full != rows:                                       # ragged last row
            tail = src[full * self.n2:]
            lo, hi = self.rank * w, min((self.rank + 1) * w, tail.numel())
            if hi > lo:
                dst[: hi - lo, full] = tail[lo:hi]
        self._handoff_to_lib()
        if self.offset % ctx.modulus != 1:
            ctx.pow_mul_dev(v, self.n1, self.rank * w, False, self.n2, self.offset, 1, self.log_n)
        ctx.ntt_batch_dev(v, self.a)
        ctx.pow_mul_dev(v, self.n1, self.rank * w, True, 0, self.omega, 1, self.log_n)
        ctx.sync()
        t = _as_torch(v).view(w, self.world, self.n1 // self.world).permute(1, 0, 2).contiguous()
        self._keep = [v, cv]
        return t

    def phase_c(self, recv):
        """recv: [world (src)][n2/G][n1/G] == [n2][k1'] -> tensor [world (dest)][n1/G][n2/G] for the second exchange."""
        ctx = self.ctx
        v = ctx.zeros(self.n2 * (self.n1 // self.world))
        ctx.sync()
        _as_torch(v).view(self.n1 // self.world, self.n2).copy_(recv.reshape(self.n2, self.n1 // self.world).t())   # [k1'][n2]
        self._handoff_to_lib()
        ctx.ntt_batch_dev(v, self.b)
        ctx.sync()
        t = _as_torch(v).view(self.n1 // self.world, self.world, self.n2 // self.world).permute(1, 0, 2).contiguous()
        self._keep = [v]
        return t

    def finish(self, recv):
        """recv: [world (src)][n1/G][n2/G] == [k1][k2'] -> Vec with this rank's natural-order block."""
        out = self.ctx.zeros(self.n1 * (self.n2 // self.world))
        self.ctx.sync()
        _as_torch(out).view(self.n2 // self.world, self.n1).copy_(recv.reshape(self.n1, self.n2 // self.world).t())  # p = k1 + N1*k2'
        self._handoff_to_lib()
        self._keep = []
        return out


def _all_to_all(t, group=None):
    import torch
    dist = _dist()
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return t.clone()
    out = torch.empty_like(t)
    dist.all_to_all_single(out.view(-1), t.view(-1), group=group)
    return out


def four_step_lde(sp, ctx, coeffs: np.ndarray, log_n: int, offset: int, rank: int, world: int, group=None):
    """This rank's contiguous natural-order block of the evaluations of `coeffs` on offset*<w_{2^log_n}>."""
    fs = FourStepLDE(sp, ctx, log_n, offset, rank, world)
    r1 = _all_to_all(fs.phase_a(coeffs), group)
    r2 = _all_to_all(fs.phase_c(r1), group)
    return fs.finish(r2)


def four_step_lde_emulated(sp, ctx, coeffs: np.ndarray, log_n: int, offset: int, world: int) -> list:
    """All `world` ranks of the flow on ONE GPU (the exchanges happen in memory): exercises the multi-rank
    index algebra and every local kernel without needing `world` devices."""
    import torch
    ranks = [FourStepLDE(sp, ctx, log_n, offset, r, world) for r in range(world)]
    send = [fs.phase_a(coeffs).clone() for fs in ranks]
    recv = [torch.stack([send[src][dst] for src in range(world)]) for dst in range(world)]
    send2 = [fs.phase_c(recv[r]).clone() for r, fs in enumerate(ranks)]
    recv2 = [torch.stack([send2[src][dst] for src in range(world)]) for dst in range(world)]
    return [fs.finish(recv2[r]) for r, fs in enumerate(ranks)]


# ------------------------------------------------------------------------------------------------ cfg5: FRI commit with a sharded layer 0
class MultiGpuFri:
    """What `fri_commit_multi` leaves behind for `decommit_fri_multi`: the owner's leaf range and subtree on every
    rank, the subtree roots, and (rank 0) the FRIProof whose layer 0 is the gathered evaluations."""

    def __init__(self, log_n, block, subtree, subtree_roots, proof):
        self.log_n, self.block, self.subtree, self.subtree_roots, self.proof = log_n, block, subtree, subtree_roots, proof


def _bcast_int(value: int, group=None) -> int:
    import torch
    dist = _dist()
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return value
    t = torch.tensor([value], dtype=torch.int64, device=_comm_device(group))
    dist.broadcast(t, src=0, group=group)
    return int(t.item())


def fri_commit_multi(sp, ctx, coeffs, log_n: int, offset: int, channel, rank: int, world: int, group=None,
                     p2p: Optional["FourStepP2P"] = None) -> MultiGpuFri:
    """fri_commit (reference src/fri/fri_commit.rs:72-122) with layer 0 spread over `world` GPUs:
    four-step LDE -> each rank hashes its contiguous leaf range -> subtree roots gathered -> the evaluations are
    all-gathered and rank 0 runs the (unpartitioned) fold/commit loop against the channel.  The transcript is the
    single-GPU one, byte for byte.  `channel` is only used on rank 0."""
    import torch
    dist = _dist()
    cvec = coeffs if hasattr(coeffs, "device_ptr") else ctx.upload(coeffs)
    # exchanges either over NCCL (all_to_all_single) or stored straight into peer memory by the kernels
    block = p2p.run(cvec, offset) if p2p is not None else four_step_lde(sp, ctx, cvec, log_n, offset, rank, world, group)
    subtree = sp.MerkleTree.new(ctx, block)
    root0, subs = commit_leaf_ranges(subtree.root_bytes, rank, world, group)
    ctx.sync()
    mine = _as_torch(block)
    if world > 1:
        full = torch.empty(mine.numel() * world, dtype=mine.dtype, device=mine.device)
        dist.all_gather_into_tensor(full, mine, group=group)
    else:
        full = mine
    proof = None
    if rank == 0:
        torch.cuda.current_stream().synchronize()
        layer0 = ctx.from_device(full.data_ptr(), full.numel())
        proof = sp.fri_begin_external(ctx, cvec, log_n, offset, layer0, root0)
        channel.send(root0.hex().encode())                                   # fri_commit.rs:86
        while proof.degree >= 1:                                             # :89
            beta = channel.receive_random_field_element()                    # :91
            channel.send(proof.fold(beta).hex().encode())                    # :94-100
        fin = proof.final_poly()
        channel.send(int(fin[0] if len(fin) else 0).to_bytes(8, "big"))      # :109-114
    return MultiGpuFri(log_n, block, subtree, subs, proof)


def feed_layer_records(channel, blob: bytes, layer_lens: Sequence[int], index: int) -> None:
    """Sends the records of `stark_fri_open_layers` in the reference's order (fri_commit.rs:145-163)."""
    off = 0
    for n in layer_lens:
        idx, depth = index % n, n.bit_length() - 1
        if n == 1:
            channel.send(blob[off:off + 8])                                  # :147-149, then falls through as written
        for _ in (idx, (idx + n // 2) % n):
            plen = 32 * depth
            channel.send(blob[off:off + 8])
            channel.send(blob[off + 8:off + 8 + plen])
            off += 8 + plen
    assert off == len(blob)


def decommit_fri_multi(sp, mp: MultiGpuFri, num_queries: int, max_index: int, channel, rank: int, world: int, group=None) -> None:
    """decommit_fri (fri_commit.rs:168-179) when layer 0's tree lives in leaf ranges on `world` GPUs: the index is
    drawn on rank 0 and broadcast; the owners of idx and idx + N/2 open their subtree, rank 0 appends the top
    levels and feeds the channel; the remaining layers are opened on rank 0."""
    n = 1 << mp.log_n
    blk = n // world
    depth_local = blk.bit_length() - 1
    for _ in range(num_queries):
        idx = channel.receive_random_int(0, max_index, True) if rank == 0 else 0
        idx = _bcast_int(idx, group)
        i0 = idx % n
        recs = []
        for which in (i0, (i0 + n // 2) % n):
            owner, local = which // blk, which % blk
            payload = np.zeros(8 + 32 * depth_local, dtype=np.uint8)
            if rank == owner:
                val = int(mp.block.download(local, 1)[0])
                payload[:] = np.frombuffer(val.to_bytes(8, "big") + mp.subtree.get_authentication_path(local), dtype=np.uint8)
            allp = all_gather_bytes(payload, group)
            recs.append((allp[owner, :8].tobytes(), allp[owner, 8:].tobytes() + top_path(mp.subtree_roots, owner)))
        if rank == 0:
            for elem, path in recs:                                          # layer 0: :156-163
                channel.send(elem)
                channel.send(path)
            lens = [mp.proof.layer_len(k) for k in range(1, mp.proof.num_layers)]
            feed_layer_records(channel, mp.proof.open([idx], first_layer=1), lens, idx)


# ------------------------------------------------------------------------------------------------ cfg5: the large FRI layers in leaf ranges
# Host protocol of csrc/multi.cu (sharded_fri_layers, open_leaf_ranges_batch) with the compute injected, so that it runs
# over a real multi-process group on CPU (gloo; tests/test_multi_gpu_cpu.py injects the oracle): every rank holds layer 0,
# replicates the folds, hashes its own leaf range of each large layer, subtree roots are gathered, rank 0 feeds the channel
# and beta travels back; the small layers stay with rank 0; a query's leaf-range openings travel in ONE exchange.
class LeafRangeLayer:
    def __init__(self, n, block, subtree, subtree_roots):
        self.n, self.block, self.subtree, self.subtree_roots = n, block, subtree, subtree_roots


def fri_commit_leaf_range_layers(engine, coeffs, log_n: int, offset: int, modulus: int, channel, rank: int, world: int,
                                 min_len: int, group=None):
    """fri_commit (fri_commit.rs:72-122).  `engine`: evaluate(coeffs, log_n, offset) -> layer 0, fold(evals, beta, offset,
    log_len) -> next layer, fold_poly(coeffs, beta) -> trimmed coefficients, tree(values) -> object with root() / path(i).
    Returns (layers hashed in leaf ranges, [(values, tree)] of the small layers on rank 0)."""
    def commit(values):
        blk = len(values) // world
        mine = values[rank * blk:(rank + 1) * blk]
        sub = engine.tree(mine)
        root, subs = commit_leaf_ranges(sub.root, rank, world, group)
        if rank == 0:
            channel.send(root.hex().encode())                                 # :86 / :100
        return LeafRangeLayer(len(values), mine, sub, subs)

    poly = np.asarray(coeffs, dtype=np.uint64)
    while len(poly) and poly[-1] == 0:
        poly = poly[:-1]
    cur = engine.evaluate(poly, log_n, offset)
    ranged, cur_log, cur_off = [commit(cur)], log_n, offset % modulus
    while len(poly) - 1 >= 1:                                                  # :89, identical on every rank
        half = (1 << cur_log) >> 1
        if half < min_len or half // world < 2:
            break
        beta = channel.receive_random_field_element() if rank == 0 else 0     # :91
        beta = _bcast_int(beta, group)
        cur = engine.fold(cur, beta, cur_off, cur_log)                         # replicated
        poly = engine.fold_poly(poly, beta)
        ranged.append(commit(cur))
        cur_log, cur_off = cur_log - 1, cur_off * cur_off % modulus
    tail = []
    if rank == 0:
        while len(poly) - 1 >= 1:
            beta = channel.receive_random_field_element()
            cur = engine.fold(cur, beta, cur_off, cur_log)
            poly = engine.fold_poly(poly, beta)
            t = engine.tree(cur)
            channel.send(t.root().hex().encode())
            tail.append((cur, t))
            cur_log, cur_off = cur_log - 1, cur_off * cur_off % modulus
        channel.send(int(poly[0] if len(poly) else 0).to_bytes(8, "big"))      # :109-114
    return ranged, tail


def decommit_fri_leaf_range_layers(ranged, tail, num_queries: int, max_index: int, channel, rank: int, world: int, group=None) -> None:
    """decommit_fri (fri_commit.rs:168-179): one broadcast of the index and ONE gather per query for every layer hashed in ranges."""
    for _ in range(num_queries):
        idx = channel.receive_random_int(0, max_index, True) if rank == 0 else 0
        idx = _bcast_int(idx, group)
        items, off = [], [0]
        for lay in ranged:
            blk = lay.n // world
            for which in (idx % lay.n, (idx % lay.n + lay.n // 2) % lay.n):
                items.append((lay, which, blk))
                off.append(off[-1] + 8 + 32 * (blk.bit_length() - 1))
        payload = np.zeros(off[-1], dtype=np.uint8)
        for t, (lay, which, blk) in enumerate(items):
            if which // blk == rank:
                rec = int(lay.block[which % blk]).to_bytes(8, "big") + lay.subtree.path(which % blk)
                payload[off[t]:off[t + 1]] = np.frombuffer(rec, dtype=np.uint8)
        allp = all_gather_bytes(payload, group)
        if rank != 0:
            continue
        for t, (lay, which, blk) in enumerate(items):                          # :156-163, layer by layer
            rec = allp[which // blk, off[t]:off[t + 1]].tobytes()
            channel.send(rec[:8])
            channel.send(rec[8:] + top_path(lay.subtree_roots, which // blk))
        for values, tree in tail:
            n = len(values)
            i = idx % n
            if n == 1:
                channel.send(int(values[i]).to_bytes(8, "big"))               # :147-149, then falls through as written
            for which in (i, (i + n // 2) % n):
                channel.send(int(values[which]).to_bytes(8, "big"))
                channel.send(tree.path(which))


# ------------------------------------------------------------------------------------------------ cfg5: the whole FibonacciSq prover
def _open_leaf_range(block, subtree, subtree_roots, which: int, blk: int, rank: int, group=None) -> tuple[bytes, bytes]:
    """(BE8(value), authentication path) of leaf `which` of a tree committed in leaf ranges: the owner opens its
    subtree, everybody learns the record, the levels above the subtree roots are appended."""
    owner, local = which // blk, which % blk
    depth_local = blk.bit_length() - 1
    payload = np.zeros(8 + 32 * depth_local, dtype=np.uint8)
    if rank == owner:
        val = int(block.download(local, 1)[0])
        payload[:] = np.frombuffer(val.to_bytes(8, "big") + subtree.get_authentication_path(local), dtype=np.uint8)
    allp = all_gather_bytes(payload, group)
    return allp[owner, :8].tobytes(), allp[owner, 8:].tobytes() + top_path(subtree_roots, owner)


def stark101_prove_multi(sp, ctx, channel, a1: int, log_trace: int, log_blowup: int, num_queries: int, rank: int, world: int,
                         group=None, timings: Optional[dict] = None) -> None:
    """The build-defined FibonacciSq prover (csrc/stark101.cu, stark101_prove) with everything the north-star partitions
    spread over `world` GPUs, and the same transcript byte for byte:
      * trace LDE: four-step NTT, every rank ends up with a contiguous range of f on the coset;
      * Merkle commitment of f: leaf-range subtrees, subtree roots gathered;
      * composition polynomial: point-wise on the rank's own range (it needs f at i, i+blowup, i+2*blowup: a halo of
        2*blowup values from the next rank);
      * Merkle commitment of CP (FRI layer 0): leaf-range subtrees again;
      * the rest of FRI (folds, smaller trees, channel) is not partitioned: rank 0, on the gathered layer 0.
    `channel` is only used on rank 0.  `timings` (optional dict) receives this rank's wall-clock seconds per phase."""
    import time
    import torch
    dist = _dist()
    multi = world > 1 and dist.is_initialized()
    t_last = [time.perf_counter()]

    def lap(name):
        if timings is not None:
            ctx.sync()
            now = time.perf_counter()
            timings[name] = timings.get(name, 0.0) + now - t_last[0]
            t_last[0] = now
    log_n = log_trace + log_blowup
    n, blow = 1 << log_n, 1 << log_blowup
    blk = n // world
    assert blk >= 2 * blow and blk % blow == 0, "stark101_prove_multi: ranges must hold at least 2*blowup points"
    w = ctx.generator
    # ---- src/trace: sequential recurrence, replicated (every rank needs all coefficients for the four-step LDE)
    f_coef, last_value = sp.stark101_trace_poly(ctx, a1, log_trace)
    lap("trace_and_interpolate")
    f_block = four_step_lde(sp, ctx, f_coef, log_n, w, rank, world, group)
    lap("four_step_lde")
    f_sub = sp.MerkleTree.new(ctx, f_block)
    f_root, f_subs = commit_leaf_ranges(f_sub.root_bytes, rank, world, group)
    lap("commit_f")
    alpha = [0, 0, 0]
    if rank == 0:
        channel.send(sp.stark101_statement(ctx.modulus, ctx.generator, log_trace, log_blowup, num_queries, last_value))
        channel.send(f_root.hex().encode())
        alpha = [channel.receive_random_field_element() for _ in range(3)]
    alpha = [_bcast_int(a, group) for a in alpha]
    # ---- src/composition on the local range: block + halo from the next rank (wraps around)
    ctx.sync()
    mine = _as_torch(f_block)
    if multi:
        heads = torch.empty(world * 2 * blow, dtype=mine.dtype, device=mine.device)
        dist.all_gather_into_tensor(heads, mine[: 2 * blow].contiguous(), group=group)
        halo = heads.view(world, 2 * blow)[(rank + 1) % world]
        ext = torch.cat([mine, halo])
        torch.cuda.current_stream().synchronize()
        f_ext = ctx.from_device(ext.data_ptr(), ext.numel())
        cp_block = sp.stark101_composition_range(ctx, f_ext, rank * blk, blk, alpha, last_value, log_trace, log_blowup)
        f_ext.free()
    else:
        cp_block = sp.stark101_composition_range(ctx, f_block, 0, n, alpha, last_value, log_trace, log_blowup)
    lap("composition")
    cp_sub = sp.MerkleTree.new(ctx, cp_block)
    cp_root, cp_subs = commit_leaf_ranges(cp_sub.root_bytes, rank, world, group)
    lap("commit_cp")
    # ---- src/fri on rank 0: layer 0 = the gathered CP evaluations, coefficients by interpolation (degree tracking)
    ctx.sync()
    cpm = _as_torch(cp_block)
    if multi:
        full = torch.empty(cpm.numel() * world, dtype=cpm.dtype, device=cpm.device)
        dist.all_gather_into_tensor(full, cpm, group=group)
    else:
        full = cpm
    proof = None
    lap("gather_layer0")
    if rank == 0:
        torch.cuda.current_stream().synchronize()
        layer0 = ctx.from_device(full.data_ptr(), full.numel())
        cp_coef = ctx.coset_interpolate_dev(layer0, w)
        lap("interpolate_cp")
        proof = sp.fri_begin_external(ctx, cp_coef, log_n, w, layer0, cp_root)
        channel.send(cp_root.hex().encode())                                 # fri_commit.rs:86
        while proof.degree >= 1:                                             # :89
            beta = channel.receive_random_field_element()                    # :91
            channel.send(proof.fold(beta).hex().encode())                    # :94-100
        fin = proof.final_poly()
        channel.send(int(fin[0] if len(fin) else 0).to_bytes(8, "big"))      # :109-114
        lap("fri_layers_rank0")
    # ---- queries: f(x), f(gx), f(g^2 x) from their owners, CP layer 0 from its owners, the other layers on rank 0
    for _ in range(num_queries):
        idx = channel.receive_random_int(0, n - 1 - 2 * blow, True) if rank == 0 else 0
        idx = _bcast_int(idx, group)
        recs = [_open_leaf_range(f_block, f_sub, f_subs, idx + k * blow, blk, rank, group) for k in range(3)]
        recs += [_open_leaf_range(cp_block, cp_sub, cp_subs, which, blk, rank, group) for which in (idx, (idx + n // 2) % n)]
        if rank == 0:
            for elem, path in recs:
                channel.send(elem)
                channel.send(path)
            lens = [proof.layer_len(k) for k in range(1, proof.num_layers)]
            feed_layer_records(channel, proof.open([idx], first_layer=1), lens, idx)
    lap("queries")
    if proof is not None:
        proof.free()
    for v in (f_sub, cp_sub):
        v.free()


# ------------------------------------------------------------------------------------------------ four-step NTT over peer memory
class FourStepP2P:
    """The four-step LDE with both exchanges written straight into peer memory (NVLink P2P stores through CUDA-IPC
    mapped pointers) by the kernels that produce the data — no pack / all-to-all / unpack passes (csrc/fourstep.cu).

    Every rank owns two peer-visible buffers of N/G elements: `rows` ([N1/G][N2], filled by every rank's phase A) and
    `block` (its natural-order slice of the result, filled by every rank's phase C), plus a small array of epoch flags.
    Handles are exchanged once.  The phases hand over ON THE DEVICE: a scatter kernel publishes its epoch to every peer
    when its stores are complete, the consumer's next phase starts with a one-warp wait kernel — no host barrier and no
    stream synchronisation inside a transform (`host_barriers=True` restores the two-barrier form for comparison).
    The product path for non-Python callers is `stark_mg_fourstep_lde` (csrc/multi.cu), which does the same with its own
    NCCL communicator; this class drives the phase entry points from Python over torch.distributed."""

    def __init__(self, sp, ctx, log_n: int, rank: int, world: int, group=None, host_barriers: bool = False, gather=None, barrier=None):
        """gather(bytes array) -> [world, n] and barrier(): the plumbing between the ranks; default torch.distributed."""
        self.sp, self.ctx, self.log_n, self.rank, self.world, self.group = sp, ctx, log_n, rank, world, group
        self.host_barriers, self.epoch = host_barriers, 0
        self._gather = gather or (lambda a: all_gather_bytes(a, group))
        self._barrier_fn = barrier
        n_loc = (1 << log_n) // world
        self.rows, h_rows = ctx.peer_alloc(n_loc)
        self.block, h_block = ctx.peer_alloc(n_loc)
        self.flags, h_flags = ctx.peer_alloc(32)
        ctx.sync()
        self._opened = []
        if world == 1:
            self.peer_rows, self.peer_blocks, self.peer_flags = [self.rows.device_ptr], [self.block.device_ptr], [self.flags.device_ptr]
        else:
            hs = self._gather(np.frombuffer(h_rows + h_block + h_flags, dtype=np.uint8))       # [world, 192]
            self.peer_rows, self.peer_blocks, self.peer_flags = [], [], []
            for r in range(world):
                if r == rank:
                    self.peer_rows.append(self.rows.device_ptr); self.peer_blocks.append(self.block.device_ptr)
                    self.peer_flags.append(self.flags.device_ptr)
                else:
                    ptrs = [ctx.peer_open(hs[r, 64 * k:64 * k + 64].tobytes()) for k in range(3)]
                    self._opened += ptrs
                    self.peer_rows.append(ptrs[0]); self.peer_blocks.append(ptrs[1]); self.peer_flags.append(ptrs[2])
            self._barrier()                               # every buffer exists and is zeroed before anyone stores into it

    def _barrier(self):
        if self._barrier_fn is not None:
            return self._barrier_fn()
        dist = _dist()
        if self.world > 1 and dist.is_initialized():
            dist.barrier(group=self.group)

    def run(self, coeffs, offset: int):
        """coeffs: device Vec (every rank holds the same coefficients).  Returns this rank's natural-order block
        (the peer-visible `block` buffer: valid until the next run; work that reads it must be enqueued on this context
        before the next run)."""
        ctx = self.ctx
        if self.host_barriers:
            self._barrier()                                   # nobody still reads `rows` / `block` from a previous run
            ctx.fourstep_phase_a(coeffs, self.log_n, offset, self.world, self.rank, self.peer_rows)
            self._barrier()                                   # every rank's rows are complete
            ctx.fourstep_phase_c(self.rows, self.log_n, self.world, self.rank, self.peer_blocks)
            self._barrier()                                   # every rank's block is complete
            return self.block
        self.epoch += 1
        ctx.fourstep_phase_a(coeffs, self.log_n, offset, self.world, self.rank, self.peer_rows, self.peer_flags, self.epoch)
        ctx.fourstep_phase_c(self.rows, self.log_n, self.world, self.rank, self.peer_blocks, self.peer_flags, self.epoch)
        ctx.fourstep_wait(self.peer_flags[self.rank], 1, self.world, self.epoch)
        return self.block

    def close(self):
        self.ctx.sync()
        self._barrier()
        for p in self._opened:
            self.ctx.peer_close(p)
        self._opened = []


def four_step_p2p_emulated(sp, ctx, coeffs, log_n: int, offset: int, world: int, flags: bool = True) -> list:
    """All ranks of FourStepP2P in ONE process on one GPU: the 'peer' pointers are plain device pointers of the
    other emulated ranks' buffers, so the kernels, the index algebra and (flags=True) the epoch hand-over run exactly as
    on `world` GPUs.  One stream: every rank's phase A is enqueued before any rank's phase C, whose wait kernel would
    otherwise spin for a producer queued behind it."""
    n_loc = (1 << log_n) // world
    rows = [ctx.peer_alloc(n_loc)[0] for _ in range(world)]
    blocks = [ctx.peer_alloc(n_loc)[0] for _ in range(world)]
    fl = [ctx.peer_alloc(32)[0] for _ in range(world)] if flags else None
    cvec = coeffs if hasattr(coeffs, "device_ptr") else ctx.upload(coeffs)
    pr, pb = [v.device_ptr for v in rows], [v.device_ptr for v in blocks]
    pf = [v.device_ptr for v in fl] if flags else None
    for epoch in ((1, 2) if flags else (0,)):           # twice with flags: the second run exercises the re-armed ticket and epoch 2
        for r in range(world):
            ctx.fourstep_phase_a(cvec, log_n, offset, world, r, pr, pf, epoch)
        for r in range(world):
            ctx.fourstep_phase_c(rows[r], log_n, world, r, pb, pf, epoch)
        if flags:
            for r in range(world):
                ctx.fourstep_wait(pf[r], 1, world, epoch)
    ctx.sync()
    for v in rows + (fl or []):
        v.free()
    return blocks
