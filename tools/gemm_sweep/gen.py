import sys
M2 = "cutlass::gemm::KernelTmaWarpSpecialized2SmFastFP32Sm100"
M2S = "cutlass::gemm::KernelTmaWarpSpecialized2SmFastFP32SmemSm100"
M1 = "cutlass::gemm::KernelTmaWarpSpecialized1SmFastFP32Sm100"
E2 = "cutlass::epilogue::TmaWarpSpecialized2Sm"
E1 = "cutlass::epilogue::TmaWarpSpecialized1Sm"
V = {  # name: (LA, LB, bands, tile, cluster, main, epi, elc, promo)
    "w0_smem_p2": ("RowMajor", "RowMajor", 3, "_256,_128,_32", "_2,_1,_1", M2S, E2, "void", 2),
    "w1_k64_p2": ("RowMajor", "RowMajor", 3, "_256,_128,_64", "_2,_1,_1", M2, E2, "void", 2),
    "w2_k64_p4": ("RowMajor", "RowMajor", 3, "_256,_128,_64", "_2,_1,_1", M2, E2, "void", 4),
    "w3_smem_k64_p4": ("RowMajor", "RowMajor", 3, "_256,_128,_64", "_2,_1,_1", M2S, E2, "void", 4),
    "w4_p2_b5": ("RowMajor", "RowMajor", 5, "_256,_128,_32", "_2,_1,_1", M2, E2, "void", 2),
    "w5_p2_nt": ("RowMajor", "ColumnMajor", 3, "_256,_128,_32", "_2,_1,_1", M2, E2, "void", 2),
    "w6_k64_p4_b5": ("RowMajor", "RowMajor", 5, "_256,_128,_64", "_2,_1,_1", M2, E2, "void", 4),
    "w7_p2": ("RowMajor", "RowMajor", 3, "_256,_128,_32", "_2,_1,_1", M2, E2, "void", 2),
}
tpl = open("variant.cu.in").read()
for name, (la, lb, bands, tile, cl, main, epi, elc, promo) in V.items():
    s = tpl
    for k, v in dict(LA=la, LB=lb, BANDS=bands, TILE=tile, CLUSTER=cl, MAIN=main, EPI=epi, ELC=elc,
                     PROMO=promo, NAME=name).items():
        s = s.replace(f"@{k}@", str(v))
    open(f"{name}.cu", "w").write(s)
print(" ".join(V))
