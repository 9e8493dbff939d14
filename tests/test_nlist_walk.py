"""Host-side model check of the list builder's periodic-image walk (csrc/k_nlist.cu, k_build_lists), rectangular and
triclinic boxes.

`walk()` below replays, in float32 like the kernel, how one i-block enumerates candidate columns: the unwrapped column
range, the image indices (kx, ky) of a column, and per kz the gap test between the block's bounding box *as that image
sees it* and the column, followed by the z-bin range.  The property checked on random boxes (OpenMM's reduced form up
to the extreme |bx| = |cx| = ax/2, |cy| = by/2; cutoff up to half the box), random column grids and random blocks:

  * COMPLETE: every lattice image of every point of the brick that lies within the cutoff of the block's bounding box is
    covered by an accepted candidate -- same image indices, the point's column, its z-bin inside the candidate's range;
  * the 45 image codes (kx in -2..2, ky, kz in -1..1) are ENOUGH: no image outside that range is ever within the cutoff.

The arithmetic here must be kept in step with k_build_lists; the GPU parity tests (bit-exact pair sets on triclinic
systems, tests/test_gpu_parity.py) check the kernel itself.  (Mutations of the model -- candidate range without the tilt,
27 image codes, gap test without the image offset -- are all caught by the property.)
"""
import itertools
import math

import numpy as np
import pytest

F = np.float32


def walk(g, box):
    """Accepted candidates of one block: list of (kx, ky, kz, wx, wy, zb0, zb1).  g: geometry dict; box: (lo, hi)."""
    (lox, loy, loz), (hix, hiy, hiz) = [tuple(F(v) for v in p) for p in box]
    colWx, colWy, binH, Lz = F(g["colWx"]), F(g["colWy"]), F(g["binH"]), F(g["Lz"])
    bx, cx, cy = F(g["bx"]), F(g["cx"]), F(g["cy"])
    ncx, ncy, nzb = g["ncx"], g["ncy"], g["nzb"]
    R = F(g["reach"])
    R2 = R*R
    tiltX, tiltY = abs(bx) + abs(cx), abs(cy)
    triclinic = tiltX != 0 or tiltY != 0
    uxLo, uxHi = math.floor((lox - R - tiltX)/colWx), math.floor((hix + R + tiltX)/colWx)
    uyLo, uyHi = math.floor((loy - R - tiltY)/colWy), math.floor((hiy + R + tiltY)/colWy)
    out = []
    for ux in range(uxLo, uxHi+1):
        for uy in range(uyLo, uyHi+1):
            if ux < -2*ncx or uy < -2*ncy:
                continue
            kx, ky = (ux + 2*ncx)//ncx - 2, (uy + 2*ncy)//ncy - 2
            if kx < -2 or kx > 2 or ky < -1 or ky > 1:
                continue
            wx, wy = ux - kx*ncx, uy - ky*ncy
            d2 = F(0)
            if not triclinic:
                gx = max(F(0), F(ux)*colWx - hix, lox - F(ux+1)*colWx)
                gy = max(F(0), F(uy)*colWy - hiy, loy - F(uy+1)*colWy)
                d2 = gx*gx + gy*gy
                if d2 > R2:
                    continue
            for kz in (-1, 0, 1):
                if triclinic:
                    offX, offY = F(ky)*bx + F(kz)*cx, F(kz)*cy
                    gx = max(F(0), F(ux)*colWx - (hix - offX), (lox - offX) - F(ux+1)*colWx)
                    gy = max(F(0), F(uy)*colWy - (hiy - offY), (loy - offY) - F(uy+1)*colWy)
                    d2 = gx*gx + gy*gy
                    if d2 > R2:
                        continue
                dz = F(math.sqrt(R2 - d2)) + F(1e-4)
                zlo, zhi = loz - dz, hiz + dz
                segLo, segHi = max(zlo, F(kz)*Lz) - F(kz)*Lz, min(zhi, F(kz+1)*Lz) - F(kz)*Lz
                if segHi < segLo:
                    continue
                zb0 = max(0, min(nzb-1, math.floor(segLo/binH)))
                zb1 = max(0, min(nzb-1, math.floor(segHi/binH)))
                out.append((kx, ky, kz, wx, wy, zb0, zb1))
    return out


def random_case(rng, triclinic, extreme):
    ax, by, cz = rng.uniform(2.0, 4.5, size=3)
    cutoff = rng.uniform(0.7, 1.0)*0.5*min(ax, by, cz) if not extreme else 0.5*min(ax, by, cz)
    if triclinic:
        t = rng.choice([-0.5, 0.5], size=3) if extreme else rng.uniform(-0.5, 0.5, size=3)
        bx, cx, cy = t[0]*ax, t[1]*ax, t[2]*by
    else:
        bx = cx = cy = 0.0
    ncx, ncy, nzb = int(rng.integers(1, 7)), int(rng.integers(1, 7)), int(rng.integers(1, 40))
    g = dict(ax=ax, by=by, Lz=cz, bx=bx, cx=cx, cy=cy, ncx=ncx, ncy=ncy, nzb=nzb, colWx=ax/ncx, colWy=by/ncy, binH=cz/nzb,
             reach=cutoff + 2e-4, cutoff=cutoff)
    # a block: a box inside one column (any z extent up to a third of the box)
    col = (int(rng.integers(0, ncx)), int(rng.integers(0, ncy)))
    lo = np.array([col[0]*g["colWx"], col[1]*g["colWy"], 0.0]) + rng.uniform(0, 0.5, size=3)*[g["colWx"], g["colWy"], cz]
    hi = lo + rng.uniform(0, 0.5, size=3)*[g["colWx"], g["colWy"], 0.6*cz]
    hi = np.minimum(hi, [(col[0]+1)*g["colWx"]*(1-1e-7), (col[1]+1)*g["colWy"]*(1-1e-7), cz*(1-1e-7)])
    return g, (lo, hi)


@pytest.mark.parametrize("triclinic,extreme", [(False, False), (False, True), (True, False), (True, True)])
def test_image_walk_is_complete(triclinic, extreme):
    rng = np.random.default_rng(1000 + 2*triclinic + extreme)
    for _ in range(200):
        g, (lo, hi) = random_case(rng, triclinic, extreme)
        accepted = walk(g, (lo, hi))
        index = {}
        for kx, ky, kz, wx, wy, zb0, zb1 in accepted:
            key = (kx, ky, kz, wx, wy)
            old = index.get(key)
            index[key] = (min(zb0, old[0]), max(zb1, old[1])) if old else (zb0, zb1)
        a = np.array([g["ax"], 0, 0])
        b = np.array([g["bx"], g["by"], 0])
        c = np.array([g["cx"], g["cy"], g["Lz"]])
        pts = rng.uniform(0, 1, size=(300, 3))*[g["ax"], g["by"], g["Lz"]]*(1 - 1e-9)
        images = np.array(list(itertools.product(range(-3, 4), range(-2, 3), range(-2, 3))))
        q = pts[:, None, :] + (images @ np.array([a, b, c]))[None, :, :]
        d = np.maximum(0, np.maximum(lo - q, q - hi))
        near = np.argwhere((d*d).sum(axis=2) <= g["cutoff"]**2)          # images within the cutoff of the block's box
        for ip, im in near:
            p, (kx, ky, kz) = pts[ip], (int(v) for v in images[im])
            wx, wy = int(p[0]/g["colWx"]), int(p[1]/g["colWy"])
            zb = min(g["nzb"]-1, int(p[2]/g["binH"]))
            assert abs(kx) <= 2 and abs(ky) <= 1 and abs(kz) <= 1, f"image ({kx},{ky},{kz}) outside the 45 codes: {g}"
            rng_z = index.get((kx, ky, kz, wx, wy))
            assert rng_z is not None, f"image ({kx},{ky},{kz}) of column ({wx},{wy}) not visited: {g} {lo} {hi}"
            assert rng_z[0] <= zb <= rng_z[1], f"z-bin {zb} outside {rng_z} for image ({kx},{ky},{kz}): {g} {lo} {hi}"


def test_rectangular_walk_visits_27_images_at_most():
    """Rectangular boxes never need kx = +-2 (the pre-triclinic builder had 27 codes)."""
    rng = np.random.default_rng(7)
    for _ in range(100):
        g, box = random_case(rng, False, True)
        assert all(abs(kx) <= 1 for kx, *_ in walk(g, box))
