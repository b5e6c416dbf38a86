"""Probe-volume sharding plan (host logic of vkx_probes_update_sharded, csrc/api.cu).

The z range of the grid is cut into K chunks of s*nranks slices (K = 1 today); inside chunk k rank r owns slices
[k*s*n + r*s, k*s*n + (r+1)*s). A chunk's atlas rows are contiguous, so one all-gather per atlas per chunk assembles
the next sampled atlases on every rank; the gather of frame f overlaps the primary traversal of frame f+1. Probes are independent within a frame
(reference src/shaders/irradiance.glsl:15-17,32-38: linear index and atlas row are monotone in z), so no reduction
crosses ranks and the result equals the single-GPU update bit for bit.
"""


def chunk_plan(rz: int, nranks: int, plane: int = 512):
    """(slices per rank per chunk, number of chunks); mirrors the C++ choice: one chunk per frame (the all-gather overlaps the
    next frame's primary traversal instead of later chunks of the same frame). `plane` = rx*ry probes per z-slice (unused)."""
    if rz % nranks != 0:
        raise ValueError("grid z resolution %d is not divisible by %d ranks" % (rz, nranks))
    return rz // nranks, 1


def rank_slices(rz: int, nranks: int, rank: int, plane: int = 512):
    """z-slices owned by `rank`, chunk by chunk: list of (z0, z1)."""
    s, K = chunk_plan(rz, nranks, plane)
    return [(k * s * nranks + rank * s, k * s * nranks + (rank + 1) * s) for k in range(K)]


def rank_probe_indices(resolution, nranks: int, rank: int):
    """Linear probe indices traced by `rank`, in chunk order."""
    rx, ry, rz = resolution
    plane = rx * ry
    out = []
    for z0, z1 in rank_slices(rz, nranks, rank, plane):
        out.extend(range(z0 * plane, z1 * plane))
    return out
