"""case lists shared by the CPU-emulation tests (small) and the GPU parity tests (small + full size)"""
import itertools

PERMS = list(itertools.permutations((0, 1, 2)))
RCC = ["R2CFFT_D", "CFFT_FORWARD_D", "CFFT_FORWARD_D"]
CCR = ["C2RFFT_D", "CFFT_BACKWARD_D", "CFFT_BACKWARD_D"]
RCC_S = ["R2CFFT_S", "CFFT_FORWARD_S", "CFFT_FORWARD_S"]
CCR_S = ["C2RFFT_S", "CFFT_BACKWARD_S", "CFFT_BACKWARD_S"]
CCC = ["CFFT_FORWARD_D"] * 3
CCC_B = ["CFFT_BACKWARD_D"] * 3
CCC_S = ["CFFT_FORWARD_S"] * 3

R2R_KINDS = ["DCT1", "DST1", "DCT2", "DST2", "DCT3", "DST3", "DCT4", "DST4"]

FASTCORE_CASES = [  # (type, n along the transform dimension): internal FFT length L, core size M
    ("DCT1_COMPLEX_D", 33), ("DCT1_REAL_D", 65),          # L = 2(n-1) = 64, 128: the core directly
    ("DST1_REAL_D", 31), ("DST1_COMPLEX_D", 63),          # L = 2(n+1) = 64, 128
    ("DCT2_REAL_D", 32), ("DST2_COMPLEX_D", 64), ("DCT3_COMPLEX_D", 32), ("DST3_REAL_D", 64), ("DST4_REAL_D", 32),
    ("DCT1_COMPLEX_D", 32),                               # L = 62 = 2*31: Bluestein on M = 128 (C4's 512 -> 1022 in small)
    ("DCT2_REAL_D", 25), ("DST1_REAL_S", 40), ("DCT3_COMPLEX_S", 36),
    ("CFFT_FORWARD_D", 100), ("CFFT_BACKWARD_D", 58), ("R2CFFT_D", 100), ("CFFT_FORWARD_S", 139),
]


def half(g, d=0):
    g = list(g)
    g[d] = g[d] // 2 + 1
    return tuple(g)

# tensor-load kernel (pow2_tload.cuh): 1D transforms whose input is unit-stride along another dimension than `dim`
# (grid, type, dim, mo1, mo2, variant prefix): both tile directions (u / v), the transform dimension as dim 1 or dim 2 of
# the tensor map, transposed and contiguous stores, partial tiles, one- and multi-box tiles, double and single
TLOAD_CASES = [
    ((20, 128, 6), "CFFT_FORWARD_D", 1, (0, 1, 2), (0, 1, 2), "tload<f64,M=128,P=8,transposed>"),
    ((20, 128, 6), "CFFT_BACKWARD_D", 1, (0, 1, 2), (1, 0, 2), "tload<f64,M=128,P=8,contiguous>"),
    ((13, 5, 64), "CFFT_FORWARD_D", 2, (0, 1, 2), (2, 1, 0), "tload<f64,M=64,P=8,contiguous>"),
    ((13, 5, 64), "CFFT_FORWARD_D", 2, (1, 0, 2), (1, 0, 2), "tload<f64,M=64,P=8,transposed>"),
    ((256, 9, 7), "CFFT_BACKWARD_D", 0, (2, 0, 1), (2, 0, 1), "tload<f64,M=256,P=8,transposed>"),
    ((256, 9, 7), "CFFT_FORWARD_D", 0, (1, 0, 2), (0, 1, 2), "tload<f64,M=256,P=8,contiguous>"),
    ((512, 3, 18), "R2CFFT_D", 0, (1, 2, 0), (0, 1, 2), "tload<f64,M=256,P=16,contiguous>"),
    ((128, 40, 3), "R2CFFT_D", 0, (1, 0, 2), (1, 0, 2), "tload<f64,M=64,P=16,transposed>"),
    ((6, 1024, 10), "CFFT_FORWARD_D", 1, (0, 1, 2), (1, 0, 2), "tload<f64,M=1024,P=8,contiguous>"),
    ((10, 3, 2048), "R2CFFT_D", 2, (0, 1, 2), (0, 1, 2), "tload<f64,M=1024,P="),
    ((34, 512, 2), "CFFT_FORWARD_S", 1, (0, 1, 2), (0, 1, 2), "tload<f32,M=512,P=16,transposed>"),
    ((34, 512, 2), "CFFT_BACKWARD_S", 1, (0, 2, 1), (2, 0, 1), "tload<f32,M=512,P=16,contiguous>"),
    ((2048, 20, 2), "R2CFFT_S", 0, (1, 0, 2), (0, 1, 2), "tload<f32,M=1024,P=16,contiguous>"),
    ((6, 4096, 3), "CFFT_FORWARD_S", 1, (0, 1, 2), (0, 1, 2), "tload<f32,M=4096,P="),
]
