"""Sharding of one ZRA archive across the GPUs of a box — host-side orchestration.

Frames are independent, so rank r of W owns the contiguous frame range frame_range(F, r, W):
  * decompression: no exchange at all — every rank decodes its range straight from the seek table
    (ZraCudaDecompressFrames); an optional all-gather (NCCL over NVLink) assembles the output;
  * compression: every rank compresses its range (ZraCudaCompressFrames), then ONE exchange step:
    the per-frame compressed sizes are all-gathered (equivalently an exclusive scan of the per-shard
    totals gives every shard its base offset) and the header is stitched (ZraShardBuildHeader);
  * random access: reads are routed to the owners of their frames (split at shard boundaries), served
    from the owners' shards and returned in request order (route_reads / sharded_random_access).
One process per GPU, torch.distributed for the plumbing (backend nccl on GPUs, gloo in the CPU tests).
"""
import ctypes as C

import numpy as np

from . import binding


def frame_range(frames, rank, world):
    """Contiguous, balanced: [frames*rank//world, frames*(rank+1)//world)."""
    return frames * rank // world, frames * (rank + 1) // world


def byte_range(uncompressed_size, frame_size, rank, world):
    """Uncompressed byte range of the rank's frames."""
    frames = (uncompressed_size + frame_size - 1) // frame_size
    f0, f1 = frame_range(frames, rank, world)
    return min(f0 * frame_size, uncompressed_size), min(f1 * frame_size, uncompressed_size)


def shard_header_size(frames, meta_size=0):
    return binding.lib().ZraShardHeaderSize(frames, meta_size)


def build_header(uncompressed_size, frame_size, frame_sizes, meta=b""):
    """Archive header from the compressed size of every frame (host only, no GPU)."""
    L = binding.lib()
    sizes = np.ascontiguousarray(frame_sizes, dtype=np.uint64)
    m = np.frombuffer(bytes(meta), dtype=np.uint8)
    out = np.empty(L.ZraShardHeaderSize(sizes.size, m.size), np.uint8)
    st = L.ZraShardBuildHeader(uncompressed_size, frame_size, m.ctypes.data_as(C.c_void_p) if m.size else None, m.size,
                               sizes.ctypes.data_as(C.POINTER(C.c_uint64)), sizes.size, out.ctypes.data_as(C.c_void_p), out.size)
    if st.zra != 0:
        raise binding.ZraError(st.zra, st.zstd)
    return out


def exchange_frame_sizes(local_sizes, frames, group=None):
    """The compress path's one exchange step. Every rank contributes the compressed sizes of its frames
    (in order); returns (all sizes in archive order, this shard's base offset, total compressed bytes).
    The base offset is the exclusive scan of the per-shard totals."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    counts = [frame_range(frames, r, world)[1] - frame_range(frames, r, world)[0] for r in range(world)]
    width = max(counts) if counts else 0
    mine = torch.zeros(max(width, 1), dtype=torch.int64, device=dev)
    local = np.ascontiguousarray(local_sizes, dtype=np.uint64)
    assert local.size == counts[rank], (local.size, counts[rank])
    if local.size:
        mine[: local.size] = torch.from_numpy(local.view(np.int64)).to(dev)
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    parts = [g[: counts[r]].cpu().numpy().view(np.uint64) for r, g in enumerate(gathered)]
    totals = np.array([int(p.sum()) for p in parts], dtype=np.uint64)
    base = int(totals[:rank].sum())
    return np.concatenate(parts) if parts else np.zeros(0, np.uint64), base, int(totals.sum())


def compress_shard(ctx, d_in, in_size, frame_size, level, checksum, d_out, out_capacity, stream=0):
    """ZraCudaCompressFrames on this rank's shard: returns (per-frame sizes, bytes produced)."""
    L = binding.lib()
    frames = (in_size + frame_size - 1) // frame_size
    sizes = np.zeros(max(frames, 1), np.uint64)
    produced = C.c_size_t(0)
    st = L.ZraCudaCompressFrames(ctx._c, C.c_void_p(d_in), in_size, frame_size, level, checksum, C.c_void_p(d_out), out_capacity,
                                 sizes.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(produced), C.c_void_p(stream))
    ctx._raise(st)
    return sizes[:frames], produced.value


# ---------------------------------------------------------------------------- random access across shards
def route_reads(offsets, sizes, uncompressed_size, frame_size, world):
    """Batched random access over a sharded archive: every read goes to the rank that owns
    `offset / frameSize`; a read that crosses a shard boundary is split into one piece per shard. Returns, per rank,
    (request index, byte offset inside the REQUEST, absolute offset, piece size) as int64 arrays, in request order.
    Bounds are the streaming twin's (offset + size <= uncompressedSize, source/zra.cpp:370)."""
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    sizes = np.ascontiguousarray(sizes, dtype=np.int64)
    if offsets.size and (offsets.min() < 0 or (offsets + sizes).max() > uncompressed_size):
        raise binding.ZraError(binding.StatusCode.OutOfBoundsAccess)
    bounds = np.array([byte_range(uncompressed_size, frame_size, r, world)[0] for r in range(world)] + [uncompressed_size], dtype=np.int64)
    out = []
    for r in range(world):
        lo, hi = bounds[r], bounds[r + 1]
        a = np.maximum(offsets, lo)
        b = np.minimum(offsets + sizes, hi)
        hit = np.nonzero(b > a)[0]
        out.append((hit.astype(np.int64), (a[hit] - offsets[hit]).astype(np.int64), a[hit].astype(np.int64), (b[hit] - a[hit]).astype(np.int64)))
    return out


def sharded_random_access(offsets, sizes, uncompressed_size, frame_size, serve, group=None):
    """Every rank calls this with ITS OWN batch of reads into the whole (sharded) archive and a `serve(abs_offsets,
    sizes) -> bytes` that answers reads inside its shard (on a GPU box: ZraCudaDecompressRABatch on the resident shard).
    Two exchange steps (all-to-all of the routed pieces, all-to-all of their bytes) and the results come back in the
    caller's request order as one uint8 array (reads back to back). Backend-agnostic: object collectives keep the CPU
    tests simple; a production caller would exchange the same arrays with all_to_all_single over NCCL."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    routed = route_reads(offsets, sizes, uncompressed_size, frame_size, world)
    # step 1: pieces to their owners (every rank publishes its W outgoing lists; rank r takes the r-th of each)
    me = dist.get_rank(group)
    everyone = [None] * world
    dist.all_gather_object(everyone, [(p[2], p[3]) for p in routed], group=group)
    inbox = [everyone[src][me] for src in range(world)]
    # serve every sender's pieces from the local shard
    answers = []
    for abs_off, sz in inbox:
        answers.append(bytes(serve(abs_off, sz)) if len(sz) else b"")
    # step 2: bytes back to the requesters
    everyone = [None] * world
    dist.all_gather_object(everyone, answers, group=group)
    sizes = np.ascontiguousarray(sizes, dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64) if sizes.size else np.zeros(0, np.int64)
    out = np.empty(int(sizes.sum()), np.uint8)
    for owner in range(world):
        req, inner, _, psz = routed[owner]
        blob = np.frombuffer(everyone[owner][me], dtype=np.uint8)
        cur = 0
        for k in range(req.size):
            n = int(psz[k])
            at = int(starts[req[k]] + inner[k])
            out[at: at + n] = blob[cur: cur + n]
            cur += n
    return out


# ---------------------------------------------------------------------------- the same exchange as tensors
class ShardedReads:
    """Data plane of batched random access over an archive whose frames are sharded contiguously across the ranks
    (BASELINE configs[3]): fixed-size reads are routed to the owners of their bytes with
    torch.distributed.all_to_all_single (NCCL on GPUs, gloo in the CPU tests), served there from the resident shard
    (`serve`: on a GPU, ZraCudaDecompressRABatch on the shard's own archive) and the bytes come back the same way, in the
    caller's request order. Everything stays a tensor on `device`; nothing is pickled.

    Shards cover equal byte ranges (`shard_bytes` each, a multiple of the frame size), the last one may be shorter. A
    read that crosses a shard boundary is split into one piece per shard (rare: handled by a second, tiny exchange).
    Bounds are the streaming twin's: offset + size <= uncompressedSize (source/zra.cpp:370)."""

    def __init__(self, uncompressed_size, shard_bytes, read_size, group=None, device="cpu"):
        import torch.distributed as dist

        self.U, self.S, self.R = int(uncompressed_size), int(shard_bytes), int(read_size)
        self.group, self.dev = group, device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)

    def _exchange(self, send, send_counts, width):
        """all_to_all of rows: `send` is [n, width] sorted by destination, send_counts[w] rows go to rank w."""
        import torch
        import torch.distributed as dist

        sc = torch.as_tensor(send_counts, dtype=torch.int64, device=self.dev)
        rc = torch.empty_like(sc)
        dist.all_to_all_single(rc, sc, group=self.group)
        rcl, scl = [int(x) for x in rc.tolist()], [int(x) for x in sc.tolist()]
        recv = torch.empty((sum(rcl), width), dtype=send.dtype, device=self.dev)
        dist.all_to_all_single(recv.view(-1), send.reshape(-1), output_split_sizes=[c * width for c in rcl],
                               input_split_sizes=[c * width for c in scl], group=self.group)
        return recv, rcl

    def read(self, offsets, serve):
        """offsets: int64 tensor [n] on `device` (absolute offsets into the whole archive). serve(abs_offsets, sizes) ->
        uint8 tensor with the pieces back to back, for pieces inside this rank's shard (tensors on `device`).
        Returns a uint8 tensor [n, read_size]."""
        import torch

        n, R, S, W = int(offsets.numel()), self.R, self.S, self.world
        if n and (int(offsets.min()) < 0 or int(offsets.max()) + R > self.U):
            raise binding.ZraError(binding.StatusCode.OutOfBoundsAccess)
        out = torch.empty((n, R), dtype=torch.uint8, device=self.dev)
        owner = torch.div(offsets, S, rounding_mode="floor")
        crossing = torch.div(offsets + (R - 1), S, rounding_mode="floor") != owner
        # ---- whole reads: one row of [offset] per read, sorted by owner; the answers come back in the same order
        idx = torch.nonzero(~crossing).view(-1)
        order = torch.argsort(owner[idx], stable=True)
        idx = idx[order]
        counts = torch.bincount(owner[idx], minlength=W)[:W].tolist()
        req, rcl = self._exchange(offsets[idx].view(-1, 1), counts, 1)
        m = int(req.shape[0])
        sizes = torch.full((m,), R, dtype=torch.int64, device=self.dev)
        served = serve(req.view(-1), sizes) if m else torch.empty(0, dtype=torch.uint8, device=self.dev)
        back, _ = self._exchange(served.view(m, R), rcl, R)
        out[idx] = back
        # ---- reads that cross a shard boundary: two pieces each (a read is shorter than a shard)
        cidx = torch.nonzero(crossing).view(-1)
        k = int(cidx.numel())
        total_cross = torch.tensor([k], dtype=torch.int64, device=self.dev)
        import torch.distributed as dist
        dist.all_reduce(total_cross, group=self.group)
        if int(total_cross.item()):
            o = offsets[cidx]
            first = (owner[cidx] + 1) * S - o                      # bytes in the owner's shard
            po = torch.cat([o, o + first])                          # piece offsets
            ps = torch.cat([first, R - first])                      # piece sizes
            pw = torch.cat([owner[cidx], owner[cidx] + 1])         # piece owners
            pr = torch.cat([torch.arange(k, device=self.dev), torch.arange(k, device=self.dev)])
            pin = torch.cat([torch.zeros(k, dtype=torch.int64, device=self.dev), first])   # offset inside the read
            order = torch.argsort(pw, stable=True)
            po, ps, pw, pr, pin = po[order], ps[order], pw[order], pr[order], pin[order]
            counts = torch.bincount(pw, minlength=W)[:W].tolist() if k else [0] * W
            req, rcl = self._exchange(torch.stack([po, ps], dim=1) if k else torch.empty((0, 2), dtype=torch.int64, device=self.dev), counts, 2)
            m = int(req.shape[0])
            served = serve(req[:, 0].contiguous(), req[:, 1].contiguous()) if m else torch.empty(0, dtype=torch.uint8, device=self.dev)
            # pieces have different sizes: pad every piece to a row of R bytes for the way back
            rows = torch.zeros((m, R), dtype=torch.uint8, device=self.dev)
            at = 0
            for i in range(m):
                s = int(req[i, 1])
                rows[i, :s] = served[at: at + s]
                at += s
            back, _ = self._exchange(rows, rcl, R)
            for i in range(int(pr.numel())):
                s, a = int(ps[i]), int(pin[i])
                out[cidx[pr[i]], a: a + s] = back[i, :s]
        return out
