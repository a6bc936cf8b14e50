"""Multi-GPU partitioning of the hot path over the GPUs of one node (one process per GPU, torch.distributed).

Every (field, depth, wavelength) point of a PSF bank and every image of a render batch is independent, so the data
path has NO collective: each rank traces its own contiguous block.  The only exchange is the optional assembly of
the full bank on every rank (PSFNet fitting, bank-mode render): one all-gather of [P/G, 2, ks, ks] float32 blocks
over NCCL / NVLink (gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_bounds(n_items, world_size):
    """Contiguous, balanced block boundaries: rank r owns [b[r], b[r+1])."""
    base, rem = divmod(int(n_items), int(world_size))
    bounds = [0]
    for r in range(world_size):
        bounds.append(bounds[-1] + base + (1 if r < rem else 0))
    return bounds


def shard_slice(n_items, rank=None, world_size=None):
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    b = shard_bounds(n_items, world_size)
    return slice(b[rank], b[rank + 1])


def gather_blocks(local, n_total, group=None):
    """All-gather per-rank blocks (first dim = this rank's share of `n_total`) into the full tensor on every rank.
    Blocks may differ by one row; they are padded to the largest so that a single all_gather_into_tensor suffices."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    bounds = shard_bounds(n_total, world)
    biggest = max(bounds[r + 1] - bounds[r] for r in range(world))
    pad = torch.zeros((biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    out = out.reshape((world, biggest) + tuple(local.shape[1:]))
    return torch.cat([out[r, : bounds[r + 1] - bounds[r]] for r in range(world)], dim=0)


def dealt_groups(n_groups, rank=None, world_size=None):
    """Round-robin dealing of equal-sized groups (the depth slabs of a PSF bank): rank r owns groups r, r + G, r + 2G, ...
    Near and far slabs cost differently (far object points take the strict first surface), so contiguous blocks of depths leave
    the rank with the far end behind; dealing gives every rank the same mix.  Requires n_groups % world_size == 0."""
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if n_groups % world_size:
        raise ValueError(f"{n_groups} groups cannot be dealt evenly to {world_size} ranks")
    return list(range(rank, n_groups, world_size))


def gather_dealt(local, n_groups, out=None, scratch=None, group=None):
    """Assemble groups dealt by `dealt_groups` on every rank, in group order: ONE all_gather_into_tensor of the per-rank
    blocks [n_groups / G * g, ...] (g = rows per group) and one device-side reorder [rank][local group] -> [group].
    `out` / `scratch` (both [n_groups * g, ...]) may be passed in to keep the call allocation-free."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    per = n_groups // world
    g = local.shape[0] // per
    tail = tuple(local.shape[1:])
    if scratch is None:
        scratch = torch.empty((n_groups * g,) + tail, dtype=local.dtype, device=local.device)
    if out is None:
        out = torch.empty_like(scratch)
    dist.all_gather_into_tensor(scratch, local.contiguous(), group=group)
    out.view((per, world, g) + tail).copy_(scratch.view((world, per, g) + tail).transpose(0, 1))
    return out


def psf_bank_sharded(lens, points, ks, spp, wvln=0.589, param_list=None, gather=True, seed=None, group=None):
    """DP PSF bank for normalised points [P, 3], sharded by points.  Returns (L, R): the full [P, ks, ks] bank on
    every rank if `gather`, else this rank's block.  All ranks draw the SAME pupil samples (same CPU seed), as the
    reference shares one sample set between all points (optics.py:483-490)."""
    if seed is not None:
        torch.manual_seed(seed)
    sl = shard_slice(points.shape[0], group and dist.get_rank(group), group and dist.get_world_size(group))
    L, R = lens.psf_dp(points[sl], ks=ks, wvln=wvln, spp=spp, param_list=param_list)
    if not gather:
        return L, R
    both = gather_blocks(torch.stack((L, R), dim=1), points.shape[0], group)
    return both[:, 0], both[:, 1]


def render_sharded(lens, img, depth, foc_dist, gather=False, group=None):
    """Spatially varying DP render of a batch [B, 3, H, W], sharded by image (no halo, no exchange)."""
    sl = shard_slice(img.shape[0], group and dist.get_rank(group), group and dist.get_world_size(group))
    out = lens.render(img[sl], depth[sl], foc_dist[sl])
    return gather_blocks(out, img.shape[0], group) if gather else out
