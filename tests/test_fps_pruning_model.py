"""Host model of the pruned FPS kernel (camliflow_b200/csrc/fps.cu: fps_pruned_kernel): Morton-ordered buckets, a cached
(max running distance, tie key) record per bucket and the bounding-box test that lets a bucket skip a round.  Checks, on
the CPU and in the kernel's own arithmetic order, the two facts the kernel's exactness rests on:
  * the box distance computed by the SAME expression never exceeds the distance of a point inside the box (rounding is
    monotonic), so a skipped bucket's running distances could not have changed;
  * with that rule the sampled indices are identical to the unpruned algorithm, ties and duplicates included.
The kernel itself is compared bit for bit with the C oracle and the reference's own kernel in tests/test_gpu_l0.py."""
import numpy as np
import pytest

f32 = np.float32


def sqdist3(dx, dy, dz):
    """camli_sqdist3's order: fma(dz, dz, fma(dx, dx, dy * dy)), every step rounded to fp32 (the products are exact in
    fp64, the sums of two fp32-range terms round once more: the same monotone steps)."""
    dx, dy, dz = (np.asarray(v, dtype=np.float64) for v in (dx, dy, dz))
    t = (dy * dy).astype(f32).astype(np.float64)
    t = (dx * dx + t).astype(f32).astype(np.float64)
    return (dz * dz + t).astype(f32)


def key_lo(i):
    rev = int("{:010b}".format(i & 1023)[::-1], 2)
    return (rev << 22) | (0x3FFFFF - i)


def fps_plain(xyz, S):
    N = len(xyz)
    pd = np.full(N, 1e10, f32)
    lo = np.array([key_lo(i) for i in range(N)], np.uint64)
    cur, out = 0, []
    for _ in range(S):
        out.append(cur)
        c = xyz[cur]
        pd = np.minimum(pd, sqdist3(xyz[:, 0] - c[0], xyz[:, 1] - c[1], xyz[:, 2] - c[2]))
        k = (pd.view(np.uint32).astype(np.uint64) << np.uint64(32)) | lo
        cur = int(0x3FFFFF - (int(k.max()) & 0x3FFFFF))
    return out


def spread7(v):
    return sum(((v >> b) & 1) << (3 * b) for b in range(7))


def fps_pruned(xyz, S, per_bucket):
    N = len(xyz)
    lo_c, hi_c = xyz.min(0), xyz.max(0)
    ext = float((hi_c - lo_c).max())
    cell = f32(127.999 / ext) if ext > 0 else f32(0)
    q = np.clip(((xyz - lo_c) * cell).astype(np.int64), 0, 127)
    codes = np.array([spread7(int(a)) | (spread7(int(b)) << 1) | (spread7(int(c)) << 2) for a, b, c in q])
    order = np.argsort(codes, kind="stable")
    buckets = [order[i:i + per_bucket] for i in range(0, N, per_bucket)]
    boxes = [(xyz[b].min(0), xyz[b].max(0)) for b in buckets]
    lo = np.array([key_lo(i) for i in range(N)], np.uint64)
    pd = np.full(N, 1e10, f32)
    rec = [(f32(1e10), 0)] * len(buckets)                 # cached (max running distance, tie key) per bucket
    cur, out, updates = 0, [], 0
    for _ in range(S):
        out.append(cur)
        c = xyz[cur]
        for w, ids in enumerate(buckets):
            blo, bhi = boxes[w]
            e = np.maximum(np.maximum(blo - c, c - bhi), f32(0)).astype(f32)
            if sqdist3(e[0], e[1], e[2]) >= rec[w][0]:     # the kernel's test: no running distance of the bucket can drop
                continue
            updates += 1
            pd[ids] = np.minimum(pd[ids], sqdist3(xyz[ids, 0] - c[0], xyz[ids, 1] - c[1], xyz[ids, 2] - c[2]))
            k = int(((pd[ids].view(np.uint32).astype(np.uint64) << np.uint64(32)) | lo[ids]).max())
            rec[w] = (np.uint32(k >> 32).view(f32), k & 0xFFFFFFFF)
        best = max(r[0] for r in rec)
        gl = max(r[1] for r in rec if r[0] == best)
        cur = 0x3FFFFF - (gl & 0x3FFFFF)
    return out, updates / (S * len(buckets))


def test_box_distance_is_a_lower_bound_in_the_kernels_arithmetic():
    rng = np.random.default_rng(0)
    for scale in (1e-3, 1.0, 50.0):
        lo = (rng.standard_normal((2000, 3)) * scale).astype(f32)
        hi = (lo + np.abs(rng.standard_normal((2000, 3)) * scale * 0.3).astype(f32)).astype(f32)
        p = (lo + (hi - lo) * rng.random((2000, 3)).astype(f32)).astype(f32)
        p = np.minimum(np.maximum(p, lo), hi)                       # inside the box in fp32
        c = (rng.standard_normal((2000, 3)) * scale * 2).astype(f32)
        e = np.maximum(np.maximum(lo - c, c - hi), f32(0)).astype(f32)
        lb = sqdist3(e[:, 0], e[:, 1], e[:, 2])
        d = sqdist3(p[:, 0] - c[:, 0], p[:, 1] - c[:, 1], p[:, 2] - c[:, 2])
        assert bool((lb <= d).all())


@pytest.mark.parametrize("name", ["uniform", "lattice_ties", "duplicates", "identical", "surface"])
def test_pruned_sampling_equals_plain_sampling(name):
    rng = np.random.default_rng(7)
    N, S, per = 1536, 400, 48
    if name == "uniform":
        xyz = rng.random((N, 3)).astype(f32)
    elif name == "lattice_ties":
        xyz = (rng.integers(0, 6, (N, 3)) / 6).astype(f32)
    elif name == "duplicates":
        xyz = rng.random((N // 3, 3)).astype(f32)[rng.integers(0, N // 3, N)]
    elif name == "identical":
        xyz, S = np.zeros((N, 3), f32), 12
    else:
        xyz = np.stack([rng.random(N) * 40 - 20, rng.random(N) * 20 - 10, 10 + rng.random(N)], 1).astype(f32)
    plain = fps_plain(xyz, S)
    pruned, active = fps_pruned(xyz, S, per)
    assert pruned == plain
    if name in ("uniform", "surface"):
        assert active < 0.5                                          # and most buckets really skip most rounds
