"""Host-side sharding of the hot path across GPUs (one process per GPU), SURVEY.md section 8(e).

The path shards without any data-path collective: every output pixel depends on an 11x11 input window only.
  * frame batches  -> contiguous blocks of frames per rank, no halo, no reduction
  * one big image  -> horizontal strips; a rank reads 5 extra rows on each INTERIOR edge (exterior edges clamp,
                      exactly like the reference's retrieve_tile, src/ssim.cpp:515-583) and contributes one double
                      partial sum; the only cross-GPU traffic is the all-reduce of that scalar.
These helpers are pure Python (no CUDA) so that they can be tested with the gloo backend on CPU."""

HALO = 5  # Gaussian radius, reference src/ssim.cpp:227


def shard_frames(n_frames, world, rank):
    """Contiguous block of frames [first, last) owned by `rank`."""
    base, extra = divmod(n_frames, world)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def strip_bounds(height, world, rank, halo=HALO):
    """Rows of one strip.  Returns (src0, src1, out_y0, out_rows):
    the rank loads image rows [src0, src1) and produces outputs for out_rows rows starting at row out_y0 OF ITS
    BUFFER (i.e. image row src0 + out_y0).  These are the (srcRows, outY0, outRows) arguments of
    ssim_cuda_compute_device()."""
    y0 = height * rank // world
    y1 = height * (rank + 1) // world
    src0 = max(0, y0 - halo)
    src1 = min(height, y1 + halo)
    return src0, src1, y0 - src0, y1 - y0


def mean_from_partials(total_sum, width, height):
    """float(sum / double(uint32(width*height))) -- the reference's last step, src/ssim.cpp:1102."""
    import numpy as np
    return np.float32(total_sum / float((width * height) & 0xFFFFFFFF))
