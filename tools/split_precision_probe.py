"""CPU emulation (numpy, exact accumulation in float64) of operand splits for the all-pairs volume, relative to the
max-abs of the exact result -- the tolerance metric of BASELINE.json (1e-4).  Used to size the next step for the
power-capped build kernel: one full-rate fp16 pass for hi x hi plus the two cross terms in
8-bit e4m3 (scaled by 2^12), i.e. two pass-equivalents of tensor work instead of the three of bf16x3.
    python tools/split_precision_probe.py"""
import numpy as np


def round_sig(x, mant_bits, emin, vmax):
    """Round to 1 + mant_bits significant bits, subnormals below 2^emin, saturation at vmax."""
    x = np.asarray(x, dtype=np.float64)
    out = np.zeros_like(x)
    nz = x != 0
    e = np.maximum(np.floor(np.log2(np.abs(x[nz]))), emin)
    q = 2.0 ** (e - mant_bits)
    out[nz] = np.clip(np.round(x[nz] / q) * q, -vmax, vmax)
    return out


def e4m3(x):
    return round_sig(x, 3, -6, 448.0)


def bf16(x):
    xi = np.asarray(x, dtype=np.float32).copy().view(np.uint32)
    return ((xi + 0x7FFF + ((xi >> 16) & 1)) & 0xFFFF0000).view(np.float32).astype(np.float64)


def main():
    rs = np.random.RandomState(0)
    C, Q, S = 256, 600, 2.0 ** 12
    for name, (mu, sd) in {"N(0, 0.75^2)": (0.0, 0.75), "mean-heavy N(1, 1.45^2)": (1.0, 1.45)}.items():
        a = (mu + sd * rs.standard_normal((C, Q))).astype(np.float32)
        b = (mu + sd * rs.standard_normal((C, Q))).astype(np.float32)
        exact = a.astype(np.float64).T @ b.astype(np.float64) / np.sqrt(C)
        ah, bh = a.astype(np.float16).astype(np.float64), b.astype(np.float16).astype(np.float64)
        al, bl = a - ah, b - bh
        fp16_1 = ah.T @ bh / np.sqrt(C)
        fp16_8 = (ah.T @ bh + (e4m3(ah).T @ e4m3(bl * S) + e4m3(al * S).T @ e4m3(bh)) / S) / np.sqrt(C)
        ahb, bhb = bf16(a), bf16(b)
        alb, blb = bf16(a - ahb), bf16(b - bhb)
        bf16x3 = (ahb.T @ bhb + ahb.T @ blb + alb.T @ bhb) / np.sqrt(C)
        m = np.abs(exact).max()
        print(f"{name}: fp16 single pass {np.abs(fp16_1 - exact).max() / m:.2e} | fp16 + e4m3 cross terms "
              f"{np.abs(fp16_8 - exact).max() / m:.2e} | bf16x3 {np.abs(bf16x3 - exact).max() / m:.2e}")


if __name__ == "__main__":
    main()
