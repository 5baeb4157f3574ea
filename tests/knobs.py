"""Measurement knobs for tests and tools (NOT part of the product: the library reads no environment variable).

WX_KERNEL=persistent|persistent_cta, WX_RENDER_CHUNKS=k, WX_SMEM_PAD=bytes, WX_MARCH=tolerance, WX_NVTX=1 are translated
into wx_set_option calls on a context the test fixture or a tool has just created."""
import os

from woxel_b200 import _ffi


def apply_env(ctx):
    k = os.environ.get("WX_KERNEL")
    if k:
        ctx.set_option(_ffi.WX_OPT_KERNEL, {"tiled": 0, "persistent": 1, "persistent_cta": 2}[k])
    if os.environ.get("WX_RENDER_CHUNKS"):
        ctx.set_option(_ffi.WX_OPT_RENDER_CHUNKS, int(os.environ["WX_RENDER_CHUNKS"]))
    if os.environ.get("WX_SMEM_PAD"):
        ctx.set_option(_ffi.WX_OPT_SMEM_PAD, int(os.environ["WX_SMEM_PAD"]))
    if os.environ.get("WX_MARCH") in ("tolerance", "tolerance_clip"):
        ctx.set_option(_ffi.WX_OPT_MARCH, 1 if os.environ["WX_MARCH"] == "tolerance" else 2)
    if os.environ.get("WX_LONG_FIRST") in ("0", "1"):
        ctx.set_option(_ffi.WX_OPT_LONG_FIRST, int(os.environ["WX_LONG_FIRST"]))
    if os.environ.get("WX_LONG_THRESHOLD"):
        ctx.set_option(_ffi.WX_OPT_LONG_THRESHOLD, int(os.environ["WX_LONG_THRESHOLD"]))
    if os.environ.get("WX_NVTX") == "1":
        ctx.set_option(_ffi.WX_OPT_NVTX, 1)
    return ctx
