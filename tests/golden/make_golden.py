"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/libvaq_ref.so,
compiled from /root/reference by oracle/Makefile) on seeded inputs.  Run in the build
container (the reference mount is not available on the GPU box):

    python tests/golden/make_golden.py

The fixtures pin (a) the oracle port and (b) the CUDA path on machines without the reference.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import oracle as orc  # noqa: E402
from vaq_b200 import synth, train  # noqa: E402

OUT = Path(__file__).resolve().parent


def vaq_case(name, n, d, budget, M, min_bits, max_bits, nq, k, ti_clusters, seed, sift=False):
    X = synth.sift_like(n, d, seed=seed) if sift else synth.decaying_gaussian(n, d, decay=4.0, seed=seed)
    model, XP = train.train(X, budget, M, min_bits, max_bits, kmeans_iters=8, seed=seed)
    om = orc.Model(model.L, model.bits, model.centroids)
    ref = orc.Ref()
    rv = ref.vaq(om, orc.NN_HEAP)
    codes = rv.encode(XP)
    Qraw = synth.sift_like(nq, d, seed=seed + 7) if sift else synth.decaying_gaussian(nq, d, decay=4.0, seed=seed + 7)
    Q = model.project(Qraw)
    lut = rv.create_lut(Q)
    rv.set_methods(orc.NN_HEAP)
    lab_heap, dis_heap = rv.search(Q, k)
    rv.set_methods(orc.NN_EA)
    lab_ea, dis_ea = rv.search(Q, k)
    # refine: candidates = HEAP labels, raw-space exact re-rank (VAQ.cpp:849-876)
    kr = max(1, k // 2)
    Xpad = np.pad(X, ((0, 0), (0, model.D - d))) if model.D != d else X
    Qpad = np.pad(Qraw, ((0, 0), (0, model.D - d))) if model.D != d else Qraw
    ref_lab, ref_dis = rv.refine(Qpad, lab_heap, Xpad, kr)
    out = dict(L=model.L, bits=model.bits, cent_flat=om.cent_flat, eig=model.eig, X=X.astype(np.float32), XP=XP,
               codes=codes, Qraw=Qraw, Q=Q, lut=lut, k=k, lab_heap=lab_heap, dis_heap=dis_heap, lab_ea=lab_ea,
               dis_ea=dis_ea, kr=kr, refine_lab=ref_lab, refine_dis=ref_dis)
    if ti_clusters:
        ti = rv.cluster_ti(ti_clusters, -1, False, seed=1)
        rv.set_methods(orc.NN_TI | orc.NN_EA)
        for visit in (1.0, 0.25):
            rv.set_visit(visit)
            l, dd = rv.search(Q, k)
            out[f"ti_lab_v{int(visit * 100)}"] = l
            out[f"ti_dis_v{int(visit * 100)}"] = dd
        out.update({f"ti_{k_}": v for k_, v in ti.items()})
    rv.close()
    np.savez_compressed(OUT / f"{name}.npz", **out)
    print(name, "bits", model.bits.tolist(), "N", n, "recall-free fixture written")


def hamming_case():
    ref = orc.Ref()
    out = {}
    for nbits, n, nq, k, seed in ((256, 3000, 12, 10, 3), (64, 500, 6, 5, 4), (100, 700, 5, 16, 5), (512, 400, 4, 7, 6)):
        data = synth.random_bitvectors(n, nbits, seed=seed)
        # make ties more interesting: clustered copies with a few flipped bits
        q = data[:nq].copy()
        q[:, 0] ^= np.uint64(0x5)
        tag = f"b{nbits}"
        out[f"{tag}_data"] = data
        out[f"{tag}_q"] = q
        out[f"{tag}_k"] = k
        for mname, method in (("heap", orc.QM_HEAP), ("sort", orc.QM_SORT), ("heap_ea", orc.QM_HEAP_EA), ("sort_ea", orc.QM_SORT_EA)):
            idx, dist = ref.bve_query(nbits, data, q, k, method)
            out[f"{tag}_{mname}_idx"] = idx
            out[f"{tag}_{mname}_dist"] = dist
        idx, dist = ref.bve_query(nbits, data, q, k, orc.QM_HEAP, threads=2)
        out[f"{tag}_par_idx"], out[f"{tag}_par_dist"] = idx, dist
    # the reference's own KAT inputs (test/test-bitvecengine.cpp:132-134, 213-215): glibc rand() vectors
    for nbits in (1, 32, 64):
        out[f"dummy{nbits}"] = ref.generate_dummy(nbits, 5, 1)
    np.savez_compressed(OUT / "hamming.npz", **out)
    print("hamming fixture written")


if __name__ == "__main__":
    if not orc.Ref.available():
        orc.build(ref=True)
    # all K_s >= 8: bit-exact LUT path
    vaq_case("vaq_small_a", n=3000, d=32, budget=64, M=8, min_bits=5, max_bits=10, nq=16, k=10, ti_clusters=20, seed=11)
    # some K_s < 8 (fvec_L2sqr_ny fallback, L=4) and integer-valued SIFT-like rows
    vaq_case("vaq_small_b", n=2000, d=64, budget=80, M=16, min_bits=2, max_bits=9, nq=12, k=20, ti_clusters=0, seed=12, sift=True)
    hamming_case()
