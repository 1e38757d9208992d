"""Multi-GPU plumbing: one process per GPU (torchrun), whole batches of windows sharded over ranks, weights replicated,
and ONE all-gather of the per-window metric rows at the end of the pass.  Windows share no state
(reference deploy.py:112-148), but the visual towers use batch statistics (model.py:197), so the unit that may move
between ranks is a whole batch of consecutive windows of one clip, exactly as the reference would form it.
torch.distributed is only plumbing here: NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
import numpy as np
import torch
import torch.distributed as dist


def yt_all_clip_lengths(n_clips=285, seed=1234):
    """Synthetic "YT-All-shaped" clip lengths in seconds (BASELINE config 4): 285 clips (the line count of
    meta/subsets/YT-All.test.1.lst), round(clip(lognormal(ln 300, 0.6), 20, 1800))."""
    rng = np.random.RandomState(seed)
    return np.round(np.clip(rng.lognormal(np.log(300.), 0.6, size=n_clips), 20, 1800)).astype(np.int64)


def eval_schedule(clip_lengths, batch):
    """One 0.1 s window per second of every clip at t = 0.5 + k (feeder.py:222-225, 378-379), grouped into whole
    batches of `batch` consecutive windows of the same clip; the ragged tail of a clip is dropped like the reference's
    eval loop drops its tail (feeder.py:412-419).  Returns a list of (clip, first_window, n_windows == batch)."""
    out = []
    for c, L in enumerate(clip_lengths):
        n = int(L)
        for w0 in range(0, n - batch + 1, batch):
            out.append((c, w0, batch))
    return out


def shard(items, rank, world):
    """Deal clips round-robin to ranks: every batch of clip c goes to rank c % world."""
    return [it for it in items if it[0] % world == rank]


def gather_rows(rows, ids, max_rows=None, group=None):
    """The single collective of a pass: every rank contributes (n_r, K) float32 rows and (n_r, 2) int64 ids
    (clip, window); counts differ, so each rank pads to `max_rows` (the largest per-rank count, which every rank can
    derive from the shared schedule) and writes its own count into a header row of the same buffer -- one all_gather
    in total.  Without `max_rows` a scalar all_reduce(MAX) sizes the buffer first.  Returns (rows (N, K), ids (N, 2))
    in rank order on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rows = rows.float()
    ids = torch.as_tensor(ids, dtype=torch.int64, device=rows.device).reshape(-1, 2)
    if world == 1:
        return rows, ids
    n, K = rows.shape
    # counts travel inside the payload: header row 0 = (count, 0, ...); ids are carried as two float64-exact columns
    buf_cols = K + 2
    if max_rows is None:
        nmax = torch.tensor([n], dtype=torch.int64, device=rows.device)
        dist.all_reduce(nmax, op=dist.ReduceOp.MAX, group=group)
        max_rows = int(nmax.item())
    if n > max_rows:
        raise ValueError('rank holds %d rows, more than max_rows=%d' % (n, max_rows))
    m = int(max_rows)
    buf = torch.zeros((m + 1, buf_cols), dtype=torch.float64, device=rows.device)
    buf[0, 0] = float(n)
    buf[1:n + 1, :K] = rows.double()
    buf[1:n + 1, K:] = ids.double()
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    rr, ii = [], []
    for b in out:
        k = int(b[0, 0].item())
        rr.append(b[1:k + 1, :K].float())
        ii.append(b[1:k + 1, K:].long())
    return torch.cat(rr, 0), torch.cat(ii, 0)
