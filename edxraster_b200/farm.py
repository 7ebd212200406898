"""Frame-parallel multi-GPU farm: independent camera views are dealt round-robin to
ranks, each rank renders its views on its own GPU, and finished frame buffers are gathered to rank 0.

One process per GPU over torch.distributed. NCCL (NVLink/NVSwitch) carries only the gather of finished
frames — the render path itself has no collective. The same code runs on gloo with CPU tensors, which
is how the host logic is tested without GPUs.
"""
import torch
import torch.distributed as dist


def views_of_rank(num_views, world, rank):
    """View i is rendered by rank i % world."""
    return list(range(rank, num_views, world))


def gather_frames(local_frames, num_views, dst=0):
    """Collect every rank's frames on `dst` in view order.

    local_frames: tensor [n_local, H, W, C] (or [n_local, H, W]) holding this rank's views in the
    order of views_of_rank(). Returns a tensor [num_views, ...] on dst, None elsewhere.
    """
    world, rank = dist.get_world_size(), dist.get_rank()
    per_rank = (num_views + world - 1) // world
    shape = (per_rank,) + tuple(local_frames.shape[1:])
    padded = local_frames.new_zeros(shape)
    padded[: local_frames.shape[0]] = local_frames
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst)
    if rank != dst:
        return None
    out = local_frames.new_zeros((num_views,) + tuple(local_frames.shape[1:]))
    for r in range(world):
        ids = views_of_rank(num_views, world, r)
        out[ids] = bufs[r][: len(ids)]
    return out


def render_views(renderer, mesh, views, out_frames, shaded=True):
    """Render `views` (list of (model_view, proj, raster) or PackedTransform) into out_frames[k] (device tensors).
    `renderer` is a Renderer (one frame at a time) or a FrameRing (its lanes keep several views in flight and
    share the mesh; 1.3-3x the throughput)."""
    lanes = getattr(renderer, "lanes", None)
    if len(views) and out_frames[0].is_cuda:
        torch.cuda.current_stream(out_frames[0].device).synchronize()   # the renderer's streams do not follow torch's: pending fills of out_frames must land first
    for k, v in enumerate(views):
        target = (out_frames[k].data_ptr(), 0) if shaded else (0, out_frames[k].data_ptr())
        args = v if isinstance(v, (tuple, list)) else (v,)
        if lanes is not None:
            renderer.Submit(mesh, *args, color_ptr=target[0], depth_ptr=target[1])
        else:
            renderer.SetRenderTarget(*target)
            renderer.SetTransform(*args)
            renderer.RenderMesh(mesh)
    renderer.Synchronize()
    for r in (lanes if lanes is not None else [renderer]):
        r.SetRenderTarget(0, 0)


def bin_owner_mask(width, height, part, parts, bin_px=64, device=None):
    """Boolean [H, W] mask (frame-buffer order: row 0 = bottom scanline) of the pixels whose 64x64 bin is owned
    by `part` under edx_set_screen_partition(part, parts)."""
    bins_x = (width + bin_px - 1) // bin_px
    y = torch.arange(height, device=device).flip(0)          # raster y of each frame-buffer row
    x = torch.arange(width, device=device)
    b = (y[:, None] // bin_px) * bins_x + (x[None, :] // bin_px)
    return (b % parts) == part


def composite_sort_first(frames):
    """frames: [parts, H, W, ...] full-size buffers, one per partition; returns the assembled frame."""
    parts, h, w = frames.shape[0], frames.shape[1], frames.shape[2]
    out = frames[0].clone()
    for p in range(1, parts):
        m = bin_owner_mask(w, h, p, parts, device=frames.device)
        out[m] = frames[p][m]
    return out
