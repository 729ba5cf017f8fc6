import sys
M2 = "cutlass::gemm::KernelTmaWarpSpecialized2SmFastFP32Sm100"
M2S = "cutlass::gemm::KernelTmaWarpSpecialized2SmFastFP32SmemSm100"
M1 = "cutlass::gemm::KernelTmaWarpSpecialized1SmFastFP32Sm100"
M1S = "cutlass::gemm::KernelTmaWarpSpecialized1SmFastFP32SmemSm100"
E2 = "cutlass::epilogue::TmaWarpSpecialized2Sm"
E1 = "cutlass::epilogue::TmaWarpSpecialized1Sm"
V = {  # name: (LA, LB, bands, tile, cluster, main, epi, elc, promo, acc copy atom)
    "x0_base": ("RowMajor", "RowMajor", 3, "_256,_128,_32", "_2,_1,_1", M2S, E2, "void", 2, "void"),
    "x1_ld64": ("RowMajor", "RowMajor", 3, "_256,_128,_32", "_2,_1,_1", M2S, E2, "void", 2, "cute::SM100_TMEM_LOAD_32dp32b64x"),
    "x2_ld16": ("RowMajor", "RowMajor", 3, "_256,_128,_32", "_2,_1,_1", M2S, E2, "void", 2, "cute::SM100_TMEM_LOAD_32dp32b16x"),
    "x3_c22": ("RowMajor", "RowMajor", 3, "_256,_128,_32", "_2,_2,_1", M2S, E2, "void", 2, "void"),
    "x4_1sm": ("RowMajor", "RowMajor", 3, "_128,_128,_32", "_1,_1,_1", M1S, E1, "void", 2, "void"),
    "x5_1sm_c12": ("RowMajor", "RowMajor", 3, "_128,_128,_32", "_1,_2,_1", M1S, E1, "void", 2, "void"),
    "x6_ld128": ("RowMajor", "RowMajor", 3, "_256,_128,_32", "_2,_1,_1", M2S, E2, "void", 2, "cute::SM100_TMEM_LOAD_32dp32b128x"),
    "x7_tmem_ld64": ("RowMajor", "RowMajor", 3, "_256,_128,_32", "_2,_1,_1", M2, E2, "void", 2, "cute::SM100_TMEM_LOAD_32dp32b64x"),
}
tpl = open("variant.cu.in").read()
for name, (la, lb, bands, tile, cl, main, epi, elc, promo, atom) in V.items():
    s = tpl
    for k, v in dict(LA=la, LB=lb, BANDS=bands, TILE=tile, CLUSTER=cl, MAIN=main, EPI=epi, ELC=elc,
                     PROMO=promo, NAME=name, ATOM=atom).items():
        s = s.replace(f"@{k}@", str(v))
    open(f"{name}.cu", "w").write(s)
print(" ".join(V))
