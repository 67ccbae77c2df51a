"""Multi-GPU sharding of the path: one process per GPU, torch.distributed for the plumbing.

The reference is single-device (src/tracer.cpp:13).  Its kernel shards trivially because every
sample's RNG stream depends only on (sample, global pixel id, num_samples, time) (render.cl:488,:496):

* sample sharding (BASELINE config 4): launch k (with its own time_k) goes to rank k % world; each
  rank accumulates a private canvas; the exchange step at the end is ONE reduce-scatter of the float
  canvases (every rank ends up owning the sum of a 1/world slice of the pixels), `average` with the
  total launch count on every rank's own slice in parallel, and a gather of the ARGB8 slices on the
  destination rank -- a quarter of the bytes of reducing whole canvases to one rank, and no serial
  resolve.  (`tail="reduce"` keeps the plain reduce-to-root + resolve on the root.)  Differs from
  1 GPU by FP32 summation order only.
* tile sharding (BASELINE config 5): rows are dealt in interleaved bands (srt_set_row_bands); global
  pixel ids are preserved, so each pixel is bit-identical to the 1-GPU result; every rank resolves
  its bands and the ARGB8 images are combined with a MAX-reduce (unowned pixels resolve to A=255,
  RGB=0, owned ones are >= that bytewise).

There is no data-path collective inside a launch; the collective is the exchange step that follows
the last launch.  The `renderer` argument is duck-typed (tracer.Tracer on a GPU; the CPU tests pass an
oracle-backed stand-in) so the host logic is covered with gloo, world_size 2.
"""
import numpy as np


def launch_schedule(total_launches, rank, world):
    """Launch indices of this rank: k = rank, rank + world, ... (round-robin)."""
    return list(range(rank, total_launches, world))


def band_rows(height, band_height, band_index, band_count):
    """Rows owned by a rank under interleaved row bands (mirrors srt_set_row_bands)."""
    y = np.arange(height)
    return y[(y // band_height) % band_count == band_index] if band_count > 1 else y


def _dist():
    import torch.distributed as dist
    return dist


def render_sample_sharded(renderer, scene, rank, world, total_launches=None, dst=0, reduce_fn=None,
                          tail="reduce_scatter", output=None, timings=None):
    """Sample-sharded accumulation.  Returns the ARGB8 image on rank `dst` (None elsewhere).
    tail: "reduce_scatter" (default: reduce-scatter, per-rank resolve, gather of ARGB8 slices) or "reduce"
    (sum-reduce of whole canvases to dst, resolve there); a custom reduce_fn implies "reduce".
    output: optional host buffer for the image on dst.  timings: optional dict that receives CUDA events
    bracketing the exchange step (keys "exchange_begin", "exchange_end")."""
    total = scene.launches if total_launches is None else total_launches
    _set_bands(renderer, 1, 0, 1)  # row bands are sticky on the tracer: a sample-sharded run renders every row
    renderer.clear_canvas()
    _accumulate_all(renderer, [scene.render_data(k) for k in launch_schedule(total, rank, world)])
    if world > 1 and reduce_fn is None and tail == "reduce_scatter" and slice_pixels(renderer, world):
        return reduce_scatter_resolve(renderer, total, rank, world, dst, output, timings)
    if world > 1:
        _mark(renderer, timings, "exchange_begin")
        (reduce_fn or reduce_canvas)(renderer, dst)
        _mark(renderer, timings, "exchange_end")
    if rank == dst:
        return renderer.resolve(total) if output is None else renderer.resolve(total, output)
    return None


def render_tile_sharded(renderer, scene, rank, world, band_height=1, total_launches=None, dst=0,
                        gather_fn=None, output=None, timings=None):
    """Tile-sharded accumulation (interleaved row bands).  Returns ARGB8 on rank `dst`.
    band_height = 1 deals single rows round-robin: the cost of a row varies smoothly with y, so every rank gets
    the same load to within a row (config 5 on 8 GPUs: slowest rank 16 % above the fastest with 8-row bands, 3 % with
    single rows);
    a row of 1920 pixels x num_samples items is still far more than a warp needs for coherent first bounces."""
    total = scene.launches if total_launches is None else total_launches
    _set_bands(renderer, band_height, rank, world)
    try:
        renderer.clear_canvas()
        _accumulate_all(renderer, [scene.render_data(k) for k in range(total)])
        renderer.resolve_device(total)
        if world > 1:
            _mark(renderer, timings, "exchange_begin")
            img = (gather_fn or gather_output)(renderer, rank, dst) if gather_fn or output is None else \
                gather_output(renderer, rank, dst, output)
            _mark(renderer, timings, "exchange_end")
            return img
        return renderer.read_output() if output is None else renderer.read_output(output)
    finally:
        _set_bands(renderer, 1, 0, 1)  # leave the tracer rendering full frames again


def _set_bands(renderer, band_height, band_index, band_count):
    if hasattr(renderer, "set_row_bands"):
        renderer.set_row_bands(band_height, band_index, band_count)


def _accumulate_all(renderer, render_datas):
    """All launches of a rank, batched into shared persistent kernels where the renderer can (srt_render_batch)."""
    if hasattr(renderer, "accumulate_batch"):
        if render_datas:
            renderer.accumulate_batch(render_datas)
    else:
        for rd in render_datas:
            renderer.accumulate(rd)


def canvas_tensor(renderer):
    """torch view (no copy) of the renderer's device canvas via __cuda_array_interface__."""
    import torch
    return torch.as_tensor(renderer.canvas_view(), device="cuda")


def output_tensor(renderer):
    import torch
    return torch.as_tensor(renderer.output_view(), device="cuda")


def reduce_canvas(renderer, dst=0):
    """NCCL sum-reduce of the per-GPU float canvases into rank dst, ordered after the render
    launches on the tracer's own stream."""
    import torch
    dist = _dist()
    stream = torch.cuda.ExternalStream(renderer.stream_handle())
    with torch.cuda.stream(stream):
        dist.reduce(canvas_tensor(renderer), dst=dst, op=dist.ReduceOp.SUM)


def gather_output(renderer, rank, dst=0, output=None):
    """Combine the per-rank ARGB8 band images on rank dst (MAX-reduce, see module docstring)."""
    import torch
    dist = _dist()
    stream = torch.cuda.ExternalStream(renderer.stream_handle())
    with torch.cuda.stream(stream):
        out = output_tensor(renderer)
        dist.reduce(out, dst=dst, op=dist.ReduceOp.MAX)
    if rank != dst:
        stream.synchronize()
        return None
    return renderer.read_output() if output is None else renderer.read_output(output)  # stream-ordered, synchronises


def slice_pixels(renderer, world):
    """Pixels per rank when the canvas is cut into `world` equal contiguous slices; 0 if it does not divide."""
    px = renderer.width * renderer.height
    return px // world if px % world == 0 else 0


def _mark(renderer, timings, key):
    """Record a CUDA event on the tracer's stream (GPU renderers only) for the caller to read later."""
    if timings is None or not hasattr(renderer, "stream_handle"):
        return
    import torch
    ev = torch.cuda.Event(enable_timing=True)
    ev.record(torch.cuda.ExternalStream(renderer.stream_handle()))
    timings[key] = ev


class _on_stream:
    """Run torch work on the tracer's own CUDA stream; a no-op for the CPU stand-in renderer of the gloo tests."""

    def __init__(self, renderer):
        self.ctx = None
        if hasattr(renderer, "stream_handle"):
            import torch
            self.stream = torch.cuda.ExternalStream(renderer.stream_handle())
            self.ctx = torch.cuda.stream(self.stream)

    def __enter__(self):
        if self.ctx:
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx:
            self.ctx.__exit__(*exc)

    def synchronize(self):
        if self.ctx:
            self.stream.synchronize()


def _flat(renderer, what):
    """Flat torch view of the renderer's canvas / output: device memory of a Tracer, host arrays of the stand-in."""
    import torch
    if hasattr(renderer, "canvas_view"):
        return (canvas_tensor(renderer) if what == "canvas" else output_tensor(renderer)).view(-1)
    return torch.from_numpy(renderer.canvas if what == "canvas" else renderer.output).view(-1)


def _reduce_scatter_sum(dist, mine, whole):
    if dist.get_backend() == "gloo":  # gloo has no reduce-scatter: all-reduce a copy, keep the own slice (CPU tests)
        tmp = whole.clone()
        dist.all_reduce(tmp, op=dist.ReduceOp.SUM)
        off = mine.storage_offset() - whole.storage_offset()
        mine.copy_(tmp[off:off + mine.numel()])
    else:  # in place: NCCL allows recvbuff == sendbuff + rank * count
        dist.reduce_scatter_tensor(mine, whole, op=dist.ReduceOp.SUM)


def reduce_scatter_resolve(renderer, total_launches, rank, world, dst=0, output=None, timings=None, read_back=True):
    """The exchange step of sample sharding: reduce-scatter of the float canvases (rank r receives the sum of pixel
    slice r), `average` of that slice on its owner, gather of the ARGB8 slices into dst's output buffer, read-back.
    Everything is ordered on the tracer's stream.  Leaves slice `rank` of this rank's canvas holding the reduced
    values (the other slices keep the rank's private sums)."""
    dist = _dist()
    per = slice_pixels(renderer, world)
    with _on_stream(renderer) as st:
        _mark(renderer, timings, "exchange_begin")
        canvas = _flat(renderer, "canvas")
        _reduce_scatter_sum(dist, canvas[rank * per * 4:(rank + 1) * per * 4], canvas)
        renderer.resolve_device_range(total_launches, rank * per, per)
        out = _flat(renderer, "output")
        my_out = out[rank * per * 4:(rank + 1) * per * 4]
        parts = [out[r * per * 4:(r + 1) * per * 4] for r in range(world)] if rank == dst else None
        dist.gather(my_out.clone() if rank == dst else my_out, parts, dst=dst)
        _mark(renderer, timings, "exchange_end")
    if not read_back:  # asynchronous: the image stays in dst's device output buffer
        return None
    if rank != dst:
        st.synchronize()
        return None
    return renderer.read_output() if output is None else renderer.read_output(output)


def gather_reduced_canvas(renderer, rank, world, dst=0):
    """After reduce_scatter_resolve: collect the reduced canvas slices on dst (tests / parity checks only)."""
    import torch
    dist = _dist()
    per = slice_pixels(renderer, world)
    with _on_stream(renderer) as st:
        canvas = _flat(renderer, "canvas")
        mine = canvas[rank * per * 4:(rank + 1) * per * 4].clone()
        full = torch.empty_like(canvas) if rank == dst else None
        parts = [full[r * per * 4:(r + 1) * per * 4] for r in range(world)] if rank == dst else None
        dist.gather(mine, parts, dst=dst)
    st.synchronize()
    return full.view(renderer.height, renderer.width, 4).cpu().numpy() if rank == dst else None
