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
