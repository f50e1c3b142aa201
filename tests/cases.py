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


def half(g, d=0):
    g = list(g)
    g[d] = g[d] // 2 + 1
    return tuple(g)
