#!/usr/bin/env python
"""Drop-in check, way C: the reference's OWN ``src/stage1_retriever.py``, unmodified, with
``sys.modules["faiss"]`` set to ``tristage_rag_b200.faiss_compat`` -- its add_documents / search / save_index /
load_index then run on libtristage.  Compared with what the unmodified reference returned over the CPU restatement
of FAISS for the same inputs (tests/golden/pipeline_c1.json, Stage-1 part), plus the reference's IVF branch
(first batch of more than 1000 documents) against oracle/ivf.py.

  python tools/faiss_shim_check.py              # on a B200
  python tools/faiss_shim_check.py --emulate    # on the CPU emulator build of the kernels (tests/cudasim)

Needs the reference tree (authoring container only); the encoder is the deterministic fake of oracle/fakes.py."""
import argparse
import json
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--emulate", action="store_true")
    args = ap.parse_args()
    if not os.path.isdir(REF):
        print("reference tree not present: nothing to check")
        return 0
    from oracle import fakes, flat_ip
    from oracle import ivf as oivf
    from tristage_rag_b200 import _lib, faiss_compat

    if args.emulate:
        import ctypes as C
        import subprocess

        subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cudasim"), "-j", "8"], stdout=subprocess.DEVNULL)
        L = C.CDLL(os.path.join(ROOT, "build", "cudasim", "libtristage_cudasim.so"))
        for name, (res, a) in _lib.SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, a
        _lib._lib, _lib._stream_ptr = L, (lambda device: None)
        os.environ.setdefault("HOSTSIM_SM_COUNT", "16")

    os.environ["TS_STORAGE_DTYPE"] = "fp32"                   # the reference's own storage dtype: ids must match exactly
    st = types.ModuleType("sentence_transformers")
    st.SentenceTransformer = object
    sys.modules["sentence_transformers"] = st
    sys.modules["faiss"] = faiss_compat                        # <- the whole integration
    sys.path.insert(0, REF)
    import src.stage1_retriever as ref_s1                      # the reference's file, unmodified

    assert ref_s1.faiss is faiss_compat and "tristage" not in ref_s1.__file__
    enc = fakes.FakeSentenceEncoder(768)

    def load(self):
        self.model, self.embedding_dim = enc, enc.get_sentence_embedding_dimension()

    ref_s1.Stage1Retriever._load_model = load                  # weight loading only (no weights / network here)
    with open(os.path.join(ROOT, "tests", "golden", "pipeline_c1.json")) as f:
        g = json.load(f)
    checked = 0
    for case in g["cases"]:
        tmp = tempfile.mkdtemp()
        cfg = ref_s1.Stage1Config(device="cpu", cache_dir=os.path.join(tmp, "m"), index_dir=os.path.join(tmp, "i"),
                                  top_k_candidates=case["s1_topk"], batch_size=16, enable_bm25=case["enable_bm25"],
                                  bm25_top_k=case["bm25_topk"], fusion_method=case["fusion"], use_fp16=False)
        r = ref_s1.Stage1Retriever(cfg)
        r.add_documents(list(g[case["docs"]]))
        assert type(r.faiss_index) is faiss_compat.IndexFlatIP and r.get_stats()["faiss_index_type"] == "IndexFlatIP"
        for q in case["queries"]:
            got = r.search(q["query"], case["s1_topk"])
            assert [x["doc_id"] for x in got] == [x["doc_id"] for x in q["stage1"]], (case["name"], q["query"])
            for x, y in zip(got, q["stage1"]):
                assert abs(x["score"] - y["score"]) <= 1e-5 * max(1.0, abs(y["score"]))
                assert abs(x["stage1_score"] - y["stage1_score"]) <= 1e-5 * max(1.0, abs(y["stage1_score"]))
            json.dumps(got)                                    # native scalars (the MCP server dumps them)
            checked += 1
        # persistence through faiss.write_index / faiss.read_index (:436,:463)
        r.save_index()
        r2 = ref_s1.Stage1Retriever(cfg)
        r2.load_index()
        q0 = case["queries"][0]
        assert [x["doc_id"] for x in r2.search(q0["query"], case["s1_topk"])] == [x["doc_id"] for x in q0["stage1"]]

    # the reference's IVF branch: more than 1000 documents in the first batch (:262-273), later batch only added (:313)
    rng = np.random.default_rng(0)
    words = [f"w{i}" for i in range(400)]
    docs = [" ".join(rng.choice(words, size=12)) for _ in range(1300)]
    tmp = tempfile.mkdtemp()
    cfg = ref_s1.Stage1Config(device="cpu", cache_dir=os.path.join(tmp, "m"), index_dir=os.path.join(tmp, "i"),
                              top_k_candidates=20, batch_size=64, enable_bm25=False, use_fp16=False, nlist=12, nprobe=3)
    r = ref_s1.Stage1Retriever(cfg)
    r.add_documents(docs[:1100])
    r.add_documents(docs[1100:])
    assert type(r.faiss_index) is faiss_compat.IndexIVFFlat and r.faiss_index.nprobe == 3 and r.faiss_index.ntotal == 1300
    X = r._normalize_embeddings(r._encode_batch(docs)).astype(np.float32)
    r.faiss_index._ivf.sync()                                  # (a search does this by itself; the check reads the lists first)
    a = r.faiss_index._ivf.assignments()
    cent = r.faiss_index._ivf.centroids()
    for query in ("w1 w2 w3", "w10 w399 w7 w7", docs[5]):
        got = r.search(query, 20)
        qv = r._normalize_embeddings(r._encode_batch([query])).astype(np.float32)
        lists, _ = r.faiss_index._ivf.coarse_host(qv, 3)
        rD, rI = oivf.ivf_search(X, qv, a, lists, 20)
        bad = flat_ip.check_topk(np.array([[x["score"] for x in got]], np.float32), np.array([[x["doc_id"] for x in got]]),
                                 lambda b, ids: X[ids].astype(np.float64) @ qv[b].astype(np.float64), rD[:, :len(got)], rI[:, :len(got)])
        assert not bad and len(got) == int((rI[0] >= 0).sum()), bad
        assert all(got[i]["document"] == docs[got[i]["doc_id"]] for i in range(len(got)))
        checked += 1
    assert (oivf.assign_lists(X, cent) != a).mean() < 0.01
    r.save_index()
    r2 = ref_s1.Stage1Retriever(cfg)
    r2.load_index()
    assert type(r2.faiss_index).__name__ == "IndexIVFFlat" and r2.faiss_index.nprobe == 3
    assert [x["doc_id"] for x in r2.search("w1 w2 w3", 20)] == [x["doc_id"] for x in r.search("w1 w2 w3", 20)]
    print(f"faiss shim ok: the reference's own Stage1Retriever over tristage_rag_b200.faiss_compat, {checked} queries checked"
          f"{' (emulated kernels)' if args.emulate else ''}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
