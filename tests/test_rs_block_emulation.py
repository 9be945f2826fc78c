"""CPU replay of the row-streaming tcgen05 block kernel's DATA FLOW (csrc/wdsr_rs.cuh) on the operand image the host packer
builds (csrc/wdsr_rs_pack.h, read back through b200sr_wdsr_pack_block_image -- no device needed).

What it pins without a GPU: the K-major core-matrix image layouts of w1 / w2 / w3, the A-slice table of the 3x3 (incl. the PACK
form for 17..20 reduce channels), the dy-in-N accumulation into the 5-slot OUT ring with its two wrap cases, the lane stream
(strips, halo lanes, image runs), the step walk of a CTA's row range and the stored / zero-forcing predicates.  The replay follows
the kernel statement by statement; the expected values come from oracle.port.block (the reference's Block.forward restated,
models/basic_wdsr_b.py:96-144) on the same bf16-rounded operands.  Barrier protocol and hardware semantics are covered by the GPU
tests (tests/test_gpu_wdsr.py).
"""
import ctypes

import numpy as np
import pytest
import torch

NX, NT, SPAN = 16, 5, 126
XPLANE, T2PLANE, T2SLOT, MAXG3 = 2048, 2080, 6272, 5


def bf16_round(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).bfloat16().float().numpy()


def bf16_bits(a):
    return (np.ascontiguousarray(bf16_round(a)).view(np.uint32) >> 16).astype(np.uint16)


def bits_f32(u16):
    return (u16.astype(np.uint32) << 16).view(np.float32)


def layout(m1p):
    w1 = 0
    w2 = w1 + (m1p // 8) * 512
    sbo2 = (m1p // 8) * 128
    w3 = w2 + 4 * sbo2
    sbo3 = MAXG3 * 256
    b2 = w3 + 12 * sbo3
    b3 = b2 + 128
    tab = b3 + 128
    w2f = tab + 64
    return dict(w1=w1, w2=w2, w3=w3, b2=b2, b3=b3, tab=tab, w2f=w2f, total=w2f + (m1p // 16) * 3 * 256, sbo2=sbo2, sbo3=sbo3)


def operand(buf, start, lbo, sbo, rows):
    """A / B operand of one K = 16 tcgen05.mma, SWIZZLE_NONE K-major: element (r, k) at
    start + (r // 8) * sbo + (k // 8) * lbo + (r % 8) * 16 + (k % 8) * 2   (csrc/tc5.cuh)."""
    r = np.arange(rows)[:, None]
    k = np.arange(16)[None, :]
    off = start + (r // 8) * sbo + (k // 8) * lbo + (r % 8) * 16 + (k % 8) * 2
    u16 = buf[off].astype(np.uint16) | (buf[off + 1].astype(np.uint16) << 8)
    return bits_f32(u16)


class Steps:
    def __init__(self, g0, g1, H):
        self.g, self.g1, self.H = g0, g1, H
        self.unit()

    def unit(self):
        self.strip = self.g // self.H
        self.ya = self.g - self.strip * self.H
        self.yb = min(self.H, self.ya + (self.g1 - self.g))
        self.y = self.ya - 1

    def advance(self):
        if self.y < self.yb:
            self.y += 1
            return False
        self.g += self.yb - self.ya
        self.unit()
        return True

    def stored(self):
        return self.ya <= self.y < self.yb

    def in_image(self):
        return 0 <= self.y < self.H


def count_steps(g0, g1, H):
    t, g = 0, g0
    while g < g1:
        ya = g % H
        n = min(H - ya, g1 - g)
        t += n + 2
        g += n
    return t


def lane_pixel(strip, lane, N, H, W):
    slot = strip * SPAN + lane
    if slot >= N * (W + 2):
        return None
    n, xs = divmod(slot, W + 2)
    if xs < 1 or xs > W:
        return None
    return n, xs - 1


def replay(img, m1p, nc2, pack, x_planar, N, H, W, ctas):
    """x_planar: bf16-exact float32 [N][3][H][W][8].  Returns the block output in the same layout (fp32, before the bf16 store)."""
    L = layout(m1p)
    tab = img[L["tab"]:L["tab"] + 64].view(np.int32)
    ng3, a_off, a_lbo = int(tab[0]), tab[1:1 + MAXG3], tab[1 + MAXG3:1 + 2 * MAXG3]
    b2 = img[L["b2"]:L["b2"] + 128].view(np.float32)
    b3 = img[L["b3"]:L["b3"] + 128].view(np.float32)
    nstrips = (N * (W + 2) - 2 + SPAN - 1) // SPAN
    total_rows = nstrips * H
    y_out = np.full_like(x_planar, np.nan)
    written = np.zeros(x_planar.shape[:4], dtype=np.int32)
    xbits = bf16_bits(x_planar)
    for cta in range(ctas):
        g0, g1 = cta * total_rows // ctas, (cta + 1) * total_rows // ctas
        T = count_steps(g0, g1, H)
        xs = np.zeros(NX * 3 * XPLANE + XPLANE, dtype=np.uint8)
        one = np.zeros(XPLANE // 2, dtype=np.uint16)
        one[0::8] = 0x3F80
        one[1::8] = 0x3F80
        xs[NX * 3 * XPLANE:] = one.view(np.uint8)
        t2 = np.zeros(NT * T2SLOT, dtype=np.uint8)
        OUT = np.zeros((NT, 128, 32), dtype=np.float32)
        it = Steps(g0, g1, H)
        meta = []           # per step: (strip, y, stored, in_image)
        for s in range(T):
            if s > 0:
                it.advance()
            meta.append((it.strip, it.y, it.stored(), it.in_image()))

        def e3(r):
            strip, y, stored, _ = meta[r]
            acc = OUT[r % NT].copy()
            OUT[r % NT, :, :24] = 0.0
            xslot = r % NX
            for lane in range(1, SPAN + 1):
                px = lane_pixel(strip, lane, N, H, W)
                if not stored or px is None:
                    continue
                n, x = px
                for q in range(3):
                    o = xslot * 3 * XPLANE + q * XPLANE + lane * 16
                    res = bits_f32(xs[o:o + 16].view(np.uint16))
                    y_out[n, q, y, x] = acc[lane, 8 * q:8 * q + 8] + b3[8 * q:8 * q + 8] + res
                written[n, :, y, x] += 1

        for s in range(T):
            strip, y, _, in_img = meta[s]
            slot = s % NX
            # ---- producer: one run per image overlapping the strip
            a = strip * SPAN
            S = N * (W + 2)
            b = min(a + 127, S - 1)
            if in_img:
                for k in range(32):
                    n = a // (W + 2) + k
                    lo, hi = max(n * (W + 2) + 1, a), min(n * (W + 2) + W, b)
                    if n < N and lo <= hi:
                        lane0, ln, x0 = lo - a, hi - lo + 1, lo - n * (W + 2) - 1
                        for q in range(3):
                            src = xbits[n, q, y, x0:x0 + ln].reshape(-1).view(np.uint8)
                            o = slot * 3 * XPLANE + q * XPLANE + lane0 * 16
                            xs[o:o + ln * 16] = src
            # ---- G1: D1 = [planes 0,1] . w1a + [plane 2, ONE] . w1b
            base = slot * 3 * XPLANE
            A = operand(xs, base, XPLANE, 128, 128)
            B = operand(img, L["w1"], 128, 512, m1p)
            d1 = A @ B.T
            A = operand(xs, base + 2 * XPLANE, NX * 3 * XPLANE - slot * 3 * XPLANE - 2 * XPLANE, 128, 128)
            B = operand(img, L["w1"] + 256, 128, 512, m1p)
            d1 = d1 + A @ B.T
            a2 = bf16_round(np.maximum(d1, 0.0))       # E1
            # ---- G2
            d2 = np.zeros((128, 32), dtype=np.float32)
            for j in range(m1p // 16):
                B = operand(img, L["w2"] + 256 * j, 128, L["sbo2"], 32)
                d2 += a2[:, 16 * j:16 * j + 16] @ B.T
            # ---- E2
            tb = s % NT
            okv = np.array([lane_pixel(strip, lane, N, H, W) is not None for lane in range(128)]) & in_img
            v = bf16_bits(d2[:, :8 * nc2] + b2[None, :8 * nc2]) * okv[:, None].astype(np.uint16)       # [128][8 nc2]
            ent = t2[tb * T2SLOT:(tb + 1) * T2SLOT]
            for q in range(2 if pack else nc2):
                ent[q * T2PLANE + 16:q * T2PLANE + 16 + 128 * 16] = np.ascontiguousarray(v[:, 8 * q:8 * q + 8]).view(np.uint8).reshape(-1)
            if pack:
                p4 = np.ascontiguousarray(v[:, 16:20]).view(np.uint8)                             # [128][8 bytes]
                pl = ent[2 * T2PLANE:2 * T2PLANE + 130 * 16].reshape(130, 16)
                pl[1:129, 0:8] = p4          # low half of the lane's own entry
                pl[0:128, 8:16] = p4         # high half of the entry on its left
            # ---- G3
            aslot = (s + NT - 1) % NT
            acc = np.zeros((128, 96), dtype=np.float32)
            for i in range(ng3):
                A = operand(t2, tb * T2SLOT + int(a_off[i]), int(a_lbo[i]), 128, 128)
                B = operand(img, L["w3"] + 256 * i, 128, L["sbo3"], 96)
                acc += A @ B.T
            if aslot <= NT - 3:
                for g in range(3):
                    OUT[aslot + g] += acc[:, 32 * g:32 * g + 32]
            elif aslot == NT - 2:
                OUT[NT - 2] += acc[:, 0:32]
                OUT[NT - 1] += acc[:, 32:64]
                OUT[0] += acc[:, 64:96]
            else:
                if s > 0:
                    OUT[NT - 1] += acc[:, 0:32]
                OUT[0] += acc[:, 32:64]
                OUT[1] += acc[:, 64:96]
            # ---- E3 of row s - 1 (it waits for STEP_DONE of step s)
            if s >= 1:
                e3(s - 1)
    assert (written == 1).all(), "every pixel must be stored exactly once"
    return y_out


def to_planar(t):      # (N, 24, H, W) -> [N][3][H][W][8]
    n, c, h, w = t.shape
    return np.ascontiguousarray(t.reshape(n, 3, 8, h, w).transpose(0, 1, 3, 4, 2))


def from_planar(p):
    n, _, h, w, _ = p.shape
    return np.ascontiguousarray(p.transpose(0, 1, 4, 2, 3)).reshape(n, 24, h, w)


@pytest.mark.parametrize("C,M1,M2,N,H,W,ctas", [
    (24, 144, 20, 2, 9, 13, 3),      # dense widths: PACK form, 4 MMAs per row; several units per CTA
    (24, 144, 20, 3, 6, 96, 5),      # cfg2-like width: strips hold runs of two images
    (24, 144, 24, 1, 7, 140, 2),     # full third chunk: 5 MMAs per row; one image over two strips
    (20, 100, 13, 2, 5, 30, 4),      # pruned: two chunks, M1P = 112, narrow trunk padded to 24
    (9, 91, 7, 1, 11, 17, 2),        # pruned: one chunk
    (24, 144, 17, 1, 1, 200, 1),     # single-row image (both neighbours out of the image)
])
def test_replay_matches_block(C, M1, M2, N, H, W, ctas):
    from mobilesuperresolution_b200 import build, _lib
    build.build()
    lib = _lib.lib()
    rng = np.random.default_rng(C * 1000 + M2)
    w1 = (rng.standard_normal((M1, C)) * 0.2).astype(np.float32)
    b1 = (rng.standard_normal(M1) * 0.1).astype(np.float32)
    w2 = (rng.standard_normal((M2, M1)) * 0.1).astype(np.float32)
    b2 = (rng.standard_normal(M2) * 0.1).astype(np.float32)
    w3 = (rng.standard_normal((C, M2, 3, 3)) * 0.1).astype(np.float32)
    b3 = (rng.standard_normal(C) * 0.1).astype(np.float32)
    need = ctypes.c_size_t(0)
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert lib.b200sr_wdsr_pack_block_image(C, M1, M2, ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(w3), ptr(b3), None, 0, ctypes.byref(need)) == 0
    m1p = (M1 + 15) // 16 * 16
    assert need.value == layout(m1p)["total"]
    img = np.zeros(need.value, dtype=np.uint8)
    assert lib.b200sr_wdsr_pack_block_image(C, M1, M2, ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(w3), ptr(b3), ptr(img), img.size, ctypes.byref(need)) == 0
    nc2 = 1 if M2 <= 8 else 2 if M2 <= 16 else 3
    pack = 16 < M2 <= 20
    # the reduce filter as mma.sync.m16n8k16 B fragments in the K order of tcgen05.ld.16x128b (csrc/wdsr_rh.cuh, tc5.cuh): thread (g, j) of
    # n-tile nt, k-step ks holds w2[8nt + g][16ks + {j, 4 + j}] and [16ks + {8 + j, 12 + j}]; decode the region back into the padded matrix
    frag = img[layout(m1p)["w2f"]:].view(np.uint16).reshape(m1p // 16, 3, 32, 4)
    dec = np.zeros((24, m1p), dtype=np.float32)
    for ks in range(m1p // 16):
        for nt in range(3):
            for lane in range(32):
                g, j = lane // 4, lane % 4
                for i, k in enumerate((j, j + 4, j + 8, j + 12)):
                    dec[8 * nt + g, 16 * ks + k] = bits_f32(frag[ks, nt, lane, i:i + 1])[0]
    exp = np.zeros((24, m1p), dtype=np.float32)
    exp[:M2, :M1] = bf16_round(w2)
    assert np.array_equal(dec, exp)
    x = np.zeros((N, 24, H, W), dtype=np.float32)
    x[:, :C] = bf16_round(rng.standard_normal((N, C, H, W)))
    got = from_planar(replay(img, m1p, nc2, pack, to_planar(x), N, H, W, ctas))
    # expected: the reference block on the same bf16-rounded operands, rounding t1 and t2 where the kernel does
    import torch.nn.functional as F
    t = lambda a: torch.from_numpy(a).double()
    hi = bf16_round(b1)
    b1r = hi.astype(np.float64) + bf16_round(b1 - hi).astype(np.float64)
    xt = t(x[:, :C])
    t1 = F.conv2d(xt, t(bf16_round(w1))[:, :, None, None], torch.from_numpy(b1r))
    t1 = t(bf16_round(torch.relu(t1).float().numpy()))
    t2 = F.conv2d(t1, t(bf16_round(w2))[:, :, None, None], t(b2))
    t2 = t(bf16_round(t2.float().numpy()))
    ref = xt + F.conv2d(t2, t(bf16_round(w3)), t(b3), padding=1)
    err = np.abs(got[:, :C] - ref.numpy()).max()
    # fp32 vs fp64 accumulation flips single bf16 ulps of t1 / t2 (~4e-4 each through w2 / w3); a misplaced tap or channel is O(0.3)
    assert err < 8e-3 * max(1.0, float(ref.abs().max())), err
    assert C == 24 or np.abs(got[:, C:]).max() == 0.0                                    # pad channels stay exactly zero
