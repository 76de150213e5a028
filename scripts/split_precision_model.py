"""CPU model (numpy, no GPU) of two split-precision GEMM schemes, for DESIGN.md 7.1(c): the shipped 3 x TF32 split
(x = hi + lo, hi = TF32(x), lo = x - hi read by the tensor core as TF32; D = [A_hi W_hi] + [A_lo W_hi + A_hi W_lo], fp32
accumulators, products and the sum inside one MMA k-step exact) against an fp16 split with a 2^11 scale
(hi = fp16(x), lo = fp16((x - hi) * 2^11); the correction accumulator is multiplied by 2^-11 at the end), which would run
kind::f16 MMAs at twice the TF32 rate on half the operand bytes.  A MODEL: the tensor core's internal accumulation is
taken as exact within a k-step and rounded to fp32 between k-steps; scripts/tc_precision.py measures the real TF32 path."""
import numpy as np

F = np.float32


def tf32_round(x):   # tc::split_tf32: round to nearest, ties away, 10 explicit mantissa bits
    return ((x.astype(F).view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(F)


def tf32_trunc(x):   # what kind::tf32 reads of an fp32 operand
    return (x.astype(F).view(np.uint32) & np.uint32(0xFFFFE000)).view(F)


def mma_accumulate(pairs, kstep):
    """sum over k of a[:, k] * w[:, k] for every (a, w) in pairs, exact inside a k-step, fp32 between k-steps."""
    K = pairs[0][0].shape[1]
    acc = np.zeros((pairs[0][0].shape[0], pairs[0][1].shape[0]), F)
    for k0 in range(0, K, kstep):
        part = sum(a[:, k0:k0 + kstep].astype(np.float64) @ w[:, k0:k0 + kstep].astype(np.float64).T for a, w in pairs)
        acc = (acc.astype(np.float64) + part).astype(F)
    return acc


def split_tf32(A, W):
    ah, wh = tf32_round(A), tf32_round(W)
    al, wl = tf32_trunc(A - ah), tf32_trunc(W - wh)
    big = mma_accumulate([(ah, wh)], 8)
    small = mma_accumulate([(al, wh), (ah, wl)], 8)
    return big + small


def split_fp16(A, W):
    s = F(2048.0)
    ah, wh = A.astype(np.float16).astype(F), W.astype(np.float16).astype(F)
    al, wl = ((A - ah) * s).astype(np.float16).astype(F), ((W - wh) * s).astype(np.float16).astype(F)
    big = mma_accumulate([(ah, wh)], 16)
    small = mma_accumulate([(al, wh), (ah, wl)], 16)
    return big + small / s


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    print("%-26s %-9s %-22s %-22s %-22s" % ("case", "scale", "fp32 (numpy sgemm)", "3 x TF32 split", "fp16 split, 2^11"))
    for K in (128, 288):
        for dist in ("randn", "relu", "wide range"):
            A = rng.standard_normal((128, K)).astype(F)
            if dist == "relu":
                A = np.maximum(A, 0) * 2
            if dist == "wide range":   # activations spanning 1e-6 .. 1e3: fp16 subnormals and large values
                A = (A * np.exp(rng.uniform(np.log(1e-6), np.log(1e3), A.shape))).astype(F)
            W = (rng.standard_normal((128, K)) / K ** 0.5).astype(F)
            ref = A.astype(np.float64) @ W.astype(np.float64).T
            scale = np.abs(ref).max()
            cols = []
            for out in (A @ W.T, split_tf32(A, W), split_fp16(A, W)):
                e = np.abs(out.astype(np.float64) - ref)
                cols.append("max %.1e mean %.1e" % (e.max() / scale, e.mean() / scale))
            print("%-26s %-9.2f %-22s %-22s %-22s" % ("K=%d %s" % (K, dist), scale, *cols))
