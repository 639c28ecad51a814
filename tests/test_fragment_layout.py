"""CPU restatement of the data formats of the tensor-core matrix-vector path (csrc/frame_kernel.cuh gemv_mma,
csrc/engine.cu fk_build_image_kernel): the per-CTA weight image in mma.m16n8k16 A-fragment order, the input vector as
three bf16 planes in B-fragment order, and the exactness claim behind them (x = hi + mid + lo exactly; every product
weight x plane is exact in fp32). The emulation follows the PTX fragment layouts of mma.sync.m16n8k16.row.col:
  A (16x16): a0 = (row g, k 2tg..), a1 = (row g+8, k 2tg..), a2 = (row g, k 2tg+8..), a3 = (row g+8, k 2tg+8..)
  B (16x8) : b0 = (k 2tg.., col g), b1 = (k 2tg+8.., col g)          with g = lane >> 2, tg = lane & 3.
No GPU needed: this pins the format the CUDA code and the host image builder agree on."""
import numpy as np
import pytest


def bf16_rn(x: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 (round to nearest even), returned as the uint16 bit pattern."""
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    r = (u + 0x7FFF + ((u >> 16) & 1)) >> 16
    return r.astype(np.uint16)


def bf16_to_f32(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << 16).view(np.float32)


def split3(x: np.ndarray):
    """split3 / split3x2 of frame_kernel.cuh: three bf16 planes, remainders computed in fp32 (exact)."""
    x = x.astype(np.float32)
    h = bf16_rn(x)
    r1 = (x - bf16_to_f32(h)).astype(np.float32)
    m = bf16_rn(r1)
    r2 = (r1 - bf16_to_f32(m)).astype(np.float32)
    lo = bf16_rn(r2)
    return h, m, lo


def build_image(W: np.ndarray) -> np.ndarray:
    """W: [rows][K] uint16 (bf16 bits), rows % 8 == 0, K % 16 == 0 -> flat uint16 image (fk_build_image_kernel, modes 0-2)."""
    rows, K = W.shape
    nkt, nt = K // 16, rows // 8
    npair = nt // 2
    out = []
    for p in range(npair):
        for kt in range(nkt):
            for lane in range(32):
                g, tg = lane >> 2, lane & 3
                for reg in range(4):                      # a0 a1 a2 a3
                    row = (2 * p + (reg & 1)) * 8 + g
                    col = 16 * kt + 2 * tg + (reg >> 1) * 8
                    out += [W[row, col], W[row, col + 1]]
    if nt & 1:
        for kt in range(nkt):
            for lane in range(32):
                g, tg = lane >> 2, lane & 3
                for reg in range(2):                      # a0 a2 (rows 8..15 of the operand are zero registers)
                    row = (nt - 1) * 8 + g
                    col = 16 * kt + 2 * tg + reg * 8
                    out += [W[row, col], W[row, col + 1]]
    img = np.asarray(out, np.uint16)
    assert img.size == rows * K                           # no padding: image bytes = rows * K * 2
    return img


def build_bfrag(x: np.ndarray) -> np.ndarray:
    """stage_bfrag: [kt][plane 0..2][tg][reg b0 b1][2 halves] uint16, 96 bytes per 16 columns."""
    K = x.size
    h, m, lo = split3(x)
    out = np.zeros((K // 16, 3, 4, 2, 2), np.uint16)
    for k in range(K):
        kt, kk = k >> 4, k & 15
        for pl, v in enumerate((h, m, lo)):
            out[kt, pl, (kk & 7) >> 1, kk >> 3, kk & 1] = v[k]
    return out


def mma_gemv(img: np.ndarray, bfrag: np.ndarray, rows: int, K: int) -> np.ndarray:
    """gemv_mma: every (tile pair, kt) block is one m16n8k16 MMA; columns 0..2 of B carry the planes."""
    nkt, nt = K // 16, rows // 8
    npair = nt // 2
    y = np.zeros(rows, np.float64)
    for p in range(npair + (nt & 1)):
        single = p == npair
        acc = np.zeros((16, 8), np.float64)
        for kt in range(nkt):
            A = np.zeros((16, 16), np.float64)
            B = np.zeros((16, 8), np.float64)
            base = p * nkt * 512 // 2 + kt * (128 if single else 256)          # uint16 index of the block
            for lane in range(32):
                g, tg = lane >> 2, lane & 3
                if single:
                    f = img[base + lane * 4: base + lane * 4 + 4]
                    a = [f[0:2], None, f[2:4], None]
                else:
                    f = img[base + lane * 8: base + lane * 8 + 8]
                    a = [f[0:2], f[2:4], f[4:6], f[6:8]]
                for reg, pair in enumerate(a):
                    if pair is None:
                        continue
                    r, c = g + 8 * (reg & 1), 2 * tg + 8 * (reg >> 1)
                    A[r, c:c + 2] = bf16_to_f32(pair)
                if g < 3:                                    # lanes g >= 3 hold the zero columns of B
                    for reg in range(2):
                        B[2 * tg + 8 * reg: 2 * tg + 8 * reg + 2, g] = bf16_to_f32(bfrag[kt, g, tg, reg])
            acc += A @ B
        n = 8 if single else 16
        y[p * 16: p * 16 + n] = (acc[:n, 2] + acc[:n, 1]) + acc[:n, 0]
    return y


def test_split3_is_exact():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(4096).astype(np.float32) * np.float32(10.0) ** rng.integers(-6, 6, 4096).astype(np.float32),
                        np.asarray([0.0, -0.0, 1.0, -1.0, 3.1415927, 65504.0, 1e-30, -1e30], np.float32)])
    h, m, lo = split3(x)
    back = (bf16_to_f32(h).astype(np.float64) + bf16_to_f32(m).astype(np.float64)) + bf16_to_f32(lo).astype(np.float64)
    assert np.array_equal(back.astype(np.float32), x) and np.array_equal(back, x.astype(np.float64))


@pytest.mark.parametrize("rows,K", [(16, 256), (24, 256), (8, 512), (40, 256)])
def test_fragment_image_matvec(rows, K):
    rng = np.random.default_rng(rows * 1000 + K)
    W = bf16_rn(rng.standard_normal((rows, K)).astype(np.float32))
    x = rng.standard_normal(K).astype(np.float32)
    y = mma_gemv(build_image(W), build_bfrag(x), rows, K)
    ref = bf16_to_f32(W).astype(np.float64) @ x.astype(np.float64)
    assert np.allclose(y, ref, rtol=1e-12, atol=1e-12), float(np.abs(y - ref).max())
