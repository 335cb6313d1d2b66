#!/usr/bin/env python
"""Drop-in check (way A): the UNMODIFIED reference orchestrator
(/root/reference/src/retrieval_pipeline.py) is imported with `src.stage1_retriever` / `src.stage2_rescorer`
aliased to this package's modules, runs add_documents + batch_search, and its output is compared with what
the unmodified reference classes produced for the same inputs (tests/golden/pipeline_batch.json).

  python tools/dropin_check.py              # on a B200
  python tools/dropin_check.py --emulate    # on the CPU emulator build of the kernels (tests/cudasim)

Needs the reference tree (authoring container only).  Encoders are the deterministic fakes of oracle/fakes.py
(no weights / network here), Stage 3 is the word-overlap FakeReranker -- exactly what generated the golden."""
import argparse
import json
import os
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--emulate", action="store_true")
    ap.add_argument("--storage", default="fp32")
    args = ap.parse_args()
    if not os.path.isdir(REF):
        print("reference tree not present: nothing to check")
        return 0
    from oracle import fakes
    import tristage_rag_b200.stage1_retriever as s1
    import tristage_rag_b200.stage2_rescorer as s2
    from tristage_rag_b200 import _lib

    if args.emulate:
        import ctypes as C
        import subprocess

        import numpy as np
        import torch

        subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cudasim"), "-j", "8"], stdout=subprocess.DEVNULL)
        L = C.CDLL(os.path.join(ROOT, "build", "cudasim", "libtristage_cudasim.so"))
        for name, (res, a) in _lib.SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, a
        _lib._lib, _lib._stream_ptr = L, (lambda device: None)

        def rank_desc(scores, top_k, n_cand=None, device=0):
            s = np.ascontiguousarray(scores.numpy(), np.float32)
            out_s, out_p = np.empty((s.shape[0], top_k), np.float32), np.empty((s.shape[0], top_k), np.int32)
            _lib.check(L.ts_rank_desc(device, C.c_void_p(s.ctypes.data), None, s.shape[0], s.shape[1], int(top_k),
                                      C.c_void_p(out_s.ctypes.data), C.c_void_p(out_p.ctypes.data), None))
            return torch.from_numpy(out_s), torch.from_numpy(out_p)

        _lib.rank_desc = rank_desc
        real_to = torch.Tensor.to
        torch.Tensor.to = lambda self, *a, **kw: self if any(isinstance(x, torch.device) and x.type == "cuda" for x in a) else real_to(self, *a, **kw)

    # third-party modules the reference imports but this image lacks
    st = types.ModuleType("sentence_transformers")
    st.SentenceTransformer = object
    st.CrossEncoder = object
    sys.modules["sentence_transformers"] = st
    sys.modules["faiss"] = fakes.fake_faiss_module()          # imported by nothing once the alias is in place
    # INTEGRATION.md way A: alias the two stage modules, then import the reference orchestrator unchanged
    sys.path.insert(0, REF)
    import src                                                 # noqa: F401  (the reference package)
    sys.modules["src.stage1_retriever"] = s1
    sys.modules["src.stage2_rescorer"] = s2
    import src.retrieval_pipeline as rp

    assert rp.Stage1Retriever is s1.Stage1Retriever and rp.ColBERTScorer is s2.ColBERTScorer
    enc = fakes.FakeSentenceEncoder(768)
    tok = fakes.FakeTokenizer()
    tokmodel = fakes.FakeTokenModel(tok, 128)

    def load_s1(self):
        self.model, self.embedding_dim = enc, enc.get_sentence_embedding_dimension()

    def load_s2(self):
        self.tokenizer, self.model, self.use_amp = tok, tokmodel, False

    s1.Stage1Retriever._load_model = load_s1                    # weight loading only; the encoders are outside the hot path
    s2.ColBERTScorer._load_model = load_s2
    # the orchestrator builds the stage configs from its own fields; give them this run's HBM storage dtype
    rp.Stage1Config = lambda **kw: s1.Stage1Config(**{"storage_dtype": args.storage, **kw})
    rp.Stage2Config = lambda **kw: s2.Stage2Config(**{"storage_dtype": args.storage, **kw})
    rp.AdaptiveCrossEncoderReranker = lambda cfg: fakes.FakeReranker(cfg.top_k_final)

    with open(os.path.join(ROOT, "tests", "golden", "pipeline_batch.json")) as f:
        g = json.load(f)
    tol = 1e-5 if args.storage == "fp32" else 4e-3
    checked = 0
    for case in g["cases"]:
        tmp = tempfile.mkdtemp()
        cfg = rp.PipelineConfig(device="cpu", cache_dir=os.path.join(tmp, "m"), index_dir=os.path.join(tmp, "i"),
                                log_file=os.path.join(tmp, "log.txt"), log_level="ERROR", stage1_use_fp16=False,
                                stage2_use_fp16=False, auto_cleanup=False, **case["config"])
        pipe = rp.RetrievalPipeline(config=cfg)
        pipe.initialize_stages()
        assert type(pipe.stage1) is s1.Stage1Retriever and type(pipe.stage2) is s2.ColBERTScorer
        pipe.add_documents(list(g[case["docs"]]))
        for got, ref in zip(pipe.batch_search(g["queries"]), case["results"]):
            assert got["query"] == ref["query"] and sorted(got.keys()) == ref["keys"]
            if args.storage == "fp32":
                assert [x["doc_id"] for x in got["results"]] == [x["doc_id"] for x in ref["results"]], (case["name"], ref["query"])
                assert [x["doc_id"] for x in got["stage1_results"]] == ref["stage1_ids"]
                assert [x["doc_id"] for x in got["stage2_results"]] == ref["stage2_ids"]
            by_id = {x["doc_id"]: x for x in ref["results"]}
            for x in got["results"]:
                if x["doc_id"] in by_id:
                    assert abs(x["stage2_score"] - by_id[x["doc_id"]]["stage2_score"]) <= tol * max(1.0, abs(x["stage2_score"]))
            json.dumps(got["results"])
            checked += 1
        assert pipe.performance_stats["total_queries"] == case["total_queries"]
    print(f"drop-in ok: reference RetrievalPipeline over the tristage stages, {checked} queries match the reference's own output "
          f"(storage {args.storage}{', emulated kernels' if args.emulate else ''})")
    return 0


if __name__ == "__main__":
    sys.exit(main())
