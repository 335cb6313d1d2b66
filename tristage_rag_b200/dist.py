"""Row-sharded multi-GPU wrappers: one process per GPU, ``torch.distributed``.

The reference has no parallelism at all; this is the sharding
BASELINE.json's north_star asks for:

* Stage 1 -- rank r owns corpus rows ``[lo_r, hi_r)`` (contiguous), computes a
  local exact top-k with GLOBAL ids, the ``[B, k]`` (score, id) pairs are
  all-gathered (NCCL over NVLink; ``gloo`` in the CPU tests) and every rank
  merges ``G*k -> k`` with ``ts_topk_merge``.  Exact: the global top-k is a
  subset of the union of the local top-k lists.
* Stage 2 -- the token store is partitioned by the same row ranges.  After the
  Stage-1 merge every rank holds the full candidate list, so "scatter to the
  owning shard" is a local ownership filter inside ``ts_maxsim`` (unowned ids
  score 0.0) followed by one all-reduce(SUM) of the ``[B, C]`` score matrix.

Only the exchange step is a collective; the scan/scoring kernels never wait on
another rank.

Persistence: ``save(dir)`` writes one shard file per rank
(``ts_index_save`` / ``ts_tokstore_save``, layout in ``include/tristage.h``)
plus a JSON manifest with the row range of every file; ``load(dir, ...)`` works
for ANY world size -- each rank appends the pieces of the old files that
intersect its new range (``plan_reshard`` + ``ts_*_append_file``), so a corpus
saved by 8 ranks loads on 1, 2 or 4.  This replaces ``faiss.write_index`` /
``read_index`` (``src/stage1_retriever.py:436,463``) for the sharded layout.
"""
from __future__ import annotations

import json
import os
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


# The fused peer-memory exchange is the default Stage-1 data plane on CUDA (validated on 2 and 8 B200s: tools/dist_check.py
# and the parity check of bench.py, profiles/README.md); TS_P2P=0 selects the NCCL all-gather + merge kernel, and every
# rank falls back to it together when symmetric memory cannot be set up (_agree_on_setup).
P2P_DEFAULT = True


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced partition: the first ``n_total % world`` ranks get one extra row."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def owner_of(ids: torch.Tensor, n_total: int, world: int) -> torch.Tensor:
    """Rank owning each global id under ``shard_range`` (load-balance reporting)."""
    base, rem = divmod(int(n_total), int(world))
    cut = rem * (base + 1)
    big = torch.div(ids, base + 1, rounding_mode="floor")
    small = rem + torch.div(ids - cut, max(base, 1), rounding_mode="floor")
    return torch.where(ids < cut, big, small)


# ---------------------------------------------------------------- persistence --
MANIFEST_FORMAT = "tristage-shards-1"


def shard_file_name(kind: str, rank: int, world: int) -> str:
    return f"{kind}.{rank:05d}-of-{world:05d}.tsshard"


def manifest_path(directory: str, kind: str) -> str:
    return os.path.join(directory, f"{kind}.manifest.json")


def write_manifest(directory: str, kind: str, n_total: int, world: int, **meta) -> dict:
    """Manifest of a ``world``-rank save: file name and global [lo, hi) of every shard."""
    shards = []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        shards.append({"file": shard_file_name(kind, r, world), "lo": lo, "hi": hi})
    man = {"format": MANIFEST_FORMAT, "kind": kind, "n_total": int(n_total), "world_size": int(world),
           "shards": shards, **meta}
    tmp = manifest_path(directory, kind) + ".tmp"
    with open(tmp, "w") as f:
        json.dump(man, f, indent=1)
    os.replace(tmp, manifest_path(directory, kind))
    return man


def read_manifest(directory: str, kind: str) -> dict:
    with open(manifest_path(directory, kind)) as f:
        man = json.load(f)
    if man.get("format") != MANIFEST_FORMAT or man.get("kind") != kind:
        raise ValueError(f"{manifest_path(directory, kind)}: not a {kind} manifest of format {MANIFEST_FORMAT}")
    return man


def plan_reshard(shards: Sequence[Dict], lo: int, hi: int) -> List[Tuple[str, int, int]]:
    """Pieces ``(file, first_row_in_file, n_rows)`` of saved shards that cover the global range
    [lo, hi), in ascending id order.  Raises if the saved shards leave a gap inside the range."""
    pieces, at = [], lo
    for sh in sorted(shards, key=lambda x: x["lo"]):
        a, b = max(lo, sh["lo"]), min(hi, sh["hi"])
        if a >= b:
            continue
        if a != at:
            raise ValueError(f"saved shards do not cover rows [{at}, {a})")
        pieces.append((sh["file"], a - sh["lo"], b - a))
        at = b
    if at != hi and hi > lo:
        raise ValueError(f"saved shards do not cover rows [{at}, {hi})")
    return pieces


def _rank_world(group) -> Tuple[int, int]:
    if dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def _barrier(group) -> None:
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.barrier(group=group)


class PeerExchangeUnavailable(RuntimeError):
    """Raised on EVERY rank when any rank could not set the peer-memory exchange up."""


def _agree_on_setup(setup: Callable, device: torch.device, group):
    """Run ``setup()`` (symmetric-memory allocation + rendezvous) and let the group decide together whether the
    peer-memory exchange is usable: a rank that fell back to the collective on its own would leave its peers
    spinning on flags it never writes.  Every rank either gets its state or raises PeerExchangeUnavailable."""
    state, err = None, None
    try:
        state = setup()
    except (ImportError, RuntimeError, AttributeError) as e:       # no symmetric memory on this system
        err = e
    ok = torch.tensor([0 if state is None else 1], dtype=torch.int32, device=device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if int(ok.item()) == 0:
        raise PeerExchangeUnavailable(str(err) if err is not None else "another rank could not set it up")
    return state


def gather_topk(scores: torch.Tensor, ids: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """all-gather per-rank [B, k] results into [G, B, k] (rank-major = ascending id ranges)."""
    world = dist.get_world_size(group)
    B, k = scores.shape
    # dim-0 concatenation is the layout every backend (nccl, gloo) accepts
    all_s = torch.empty((world * B, k), dtype=scores.dtype, device=scores.device)
    all_i = torch.empty((world * B, k), dtype=ids.dtype, device=ids.device)
    dist.all_gather_into_tensor(all_s, scores.contiguous(), group=group)
    dist.all_gather_into_tensor(all_i, ids.contiguous(), group=group)
    return all_s.view(world, B, k), all_i.view(world, B, k)


class IVFShard:
    """Local index of a ``ShardedIndex`` in the approximate mode: this rank's rows behind inverted lists
    (``_lib.IVF``) with the SAME centroids on every rank.  A global list is then the union of the ranks' local
    lists, every rank probes the same ``nprobe`` lists, and the merged result equals what one GPU holding all
    rows returns -- the exchange step is unchanged (all-gather or peer-memory push of the [B, k] lists)."""

    def __init__(self, ivf, nprobe: int):
        self.ivf, self.nprobe = ivf, int(nprobe)

    def set_id_base(self, base: int) -> None:
        self.ivf.base.set_id_base(base)

    def search(self, q, k: int, **kw):
        return self.ivf.search(q, k, self.nprobe, **kw)

    def search_packed(self, q, k: int, blob, **kw):
        return self.ivf.search_packed(q, k, self.nprobe, blob, **kw)

    def save(self, path: str) -> None:
        self.ivf.base.save(path)


class ShardedIndex:
    """Stage-1 index whose rows are split across the ranks of a process group."""

    def __init__(self, local_index, n_total: int, group=None,
                 merge_fn: Optional[Callable] = None):
        self.local = local_index
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n_total = int(n_total)
        self.lo, self.hi = shard_range(self.n_total, self.rank, self.world)
        self.local.set_id_base(self.lo)
        if merge_fn is None:
            from ._lib import topk_merge

            merge_fn = lambda s, i: topk_merge(s, i, device=s.device.index)  # noqa: E731
            self._packed = hasattr(local_index, "search_packed")
        else:
            self._packed = False
        self._bufs = {}
        self.merge_fn = merge_fn
        # peer-memory exchange fused into the select kernel (ts_exchange, include/tristage.h) instead of all-gather +
        # merge: TS_P2P=1 / 0 forces it on / off; by default it is used when the local index is a plain ts_index on
        # a CUDA device (every rank decides together, see _agree_on_setup, and falls back to NCCL otherwise)
        env = os.environ.get("TS_P2P", "")
        fused_ok = self._packed and self.world > 1 and hasattr(local_index, "_h") and not hasattr(local_index, "ivf")
        self._p2p = bool(fused_ok and torch.cuda.is_available() and (env not in ("", "0") or (env == "" and P2P_DEFAULT)))
        self._p2p_state = None

    def search(self, q: torch.Tensor, k: int, **kw) -> Tuple[torch.Tensor, torch.Tensor]:
        """q replicated on every rank -> identical merged (scores, ids) on every rank."""
        if self.world > 1 and self._packed and q.is_cuda:
            return self._search_packed(q, k, **kw)
        s, i = self.local.search(q, k, **kw)
        if self.world == 1:
            return s, i
        all_s, all_i = gather_topk(s, i, self.group)
        return self.merge_fn(all_s, all_i)

    def save(self, directory: str) -> None:
        """Every rank writes its shard file; rank 0 writes the manifest once all files are in place."""
        os.makedirs(directory, exist_ok=True)
        self.local.save(os.path.join(directory, shard_file_name("index", self.rank, self.world)))
        _barrier(self.group)
        if self.rank == 0:
            write_manifest(directory, "index", self.n_total, self.world)
        _barrier(self.group)

    @classmethod
    def load(cls, directory: str, device: int = 0, group=None, make_local: Optional[Callable] = None,
             merge_fn: Optional[Callable] = None) -> "ShardedIndex":
        """Load a saved corpus under the CURRENT world size (which may differ from the saving one).
        ``make_local(info, reserve_rows)`` builds the empty local index (default: ``_lib.Index`` on
        ``device`` with the dim / dtype / metric of the files)."""
        from . import _lib

        man = read_manifest(directory, "index")
        rank, world = _rank_world(group)
        lo, hi = shard_range(man["n_total"], rank, world)
        info = _lib.file_probe(os.path.join(directory, man["shards"][0]["file"]))
        if make_local is None:
            local = _lib.Index(info["dim"], _lib.DTYPE_NAMES[info["dtype"]],
                               "cosine" if info["metric"] == _lib.TS_METRIC_COSINE else "ip", device,
                               reserve_rows=hi - lo)
        else:
            local = make_local(info, hi - lo)
        for fname, first, n in plan_reshard(man["shards"], lo, hi):
            local.append_file(os.path.join(directory, fname), first, n)
        return cls(local, man["n_total"], group=group, merge_fn=merge_fn)

    def _p2p_setup(self, B: int, k: int, device: torch.device):
        """Symmetric receive buffer of this rank, mapped into every rank of the group (torch's symmetric memory does
        the handle exchange), wrapped in a ``ts_exchange`` (include/tristage.h): capacity for batches up to
        max(B, 1024) x max(k, 128)."""
        import torch.distributed._symmetric_memory as symm

        from ._lib import Exchange

        B_max, k_max = max(int(B), 1024), max(int(k), 128)
        total = Exchange.buffer_bytes(self.world, B_max, k_max)
        buf = symm.empty(total, dtype=torch.uint8, device=device)
        buf.zero_()
        hdl = symm.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
        bases = [int(x) for x in hdl.buffer_ptrs]
        torch.cuda.current_stream(device).synchronize()
        hdl.barrier()                                   # every buffer is zeroed before anybody pushes into it
        return {"cap": (B_max, k_max), "buf": buf, "hdl": hdl,
                "x": Exchange(device.index, self.rank, self.world, bases, B_max, k_max)}

    def _p2p_state_for(self, B: int, k: int, device: torch.device):
        st = self._p2p_state
        if st is None or B > st["cap"][0] or k > st["cap"][1]:
            st = self._p2p_state = _agree_on_setup(lambda: self._p2p_setup(B, k, device), device, self.group)
        return st

    def _search_p2p(self, q, k, **kw):
        """scan -> select kernel pushes the [k] rows of every query into every rank's buffer over NVLink -> wait+merge
        kernel.  Four launches in one C call, no collective-library call in the steady state."""
        return self._p2p_state_for(q.shape[0], k, q.device)["x"].search(self.local, q, k, **kw)

    def search_host(self, q, k: int, out=None, **kw):
        """numpy / pinned host buffers in and out, ONE C call per step on the peer-memory data plane (H2D, step, D2H,
        synchronise); falls back to search() + copies when the exchange is not available."""
        import numpy as np

        if self._p2p and self.world > 1:
            try:
                st = self._p2p_state_for(len(q), k, torch.device("cuda", self.local.device))
                return st["x"].search_host(self.local, q, k, out=out, **kw)
            except PeerExchangeUnavailable:
                self._p2p = False
        dev = torch.device("cuda", self.local.device)
        s, i = self.search(torch.from_numpy(np.ascontiguousarray(q, np.float32)).to(dev), k, **kw)
        if out is not None:
            out[0][...] = s.cpu().numpy()
            out[1][...] = i.cpu().numpy()
            return out
        return s.cpu().numpy(), i.cpu().numpy()

    def _search_packed(self, q, k, **kw):
        """GPU fast path: scores + ids in one buffer -> ONE all-gather -> one merge kernel."""
        from ._lib import packed_layout, topk_merge_packed

        if self._p2p:
            try:
                return self._search_p2p(q, k, **kw)
            except PeerExchangeUnavailable as e:      # decided by ALL ranks together (see _agree_on_setup): NCCL path
                import logging

                logging.getLogger(__name__).warning(f"peer-memory exchange unavailable ({e}); using all-gather")
                self._p2p = False
        B = q.shape[0]
        _, nbytes = packed_layout(B, k)
        key = (B, k)
        if self._bufs.get("key") != key:
            self._bufs = {"key": key,
                          "mine": torch.empty(nbytes, dtype=torch.uint8, device=q.device),
                          "all": torch.empty(nbytes * self.world, dtype=torch.uint8, device=q.device)}
        mine, allb = self._bufs["mine"], self._bufs["all"]
        self.local.search_packed(q, k, mine, **kw)
        dist.all_gather_into_tensor(allb, mine, group=self.group)
        return topk_merge_packed(allb, self.world, B, k, device=q.device.index)


class ShardedTokStore:
    """Stage-2 token store partitioned by the same doc-id ranges."""

    def __init__(self, local_store, n_total: int, group=None):
        self.local = local_store
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n_total = int(n_total)
        self.lo, self.hi = shard_range(self.n_total, self.rank, self.world)
        self.local.set_id_base(self.lo)

    def save(self, directory: str) -> None:
        os.makedirs(directory, exist_ok=True)
        self.local.save(os.path.join(directory, shard_file_name("tokstore", self.rank, self.world)))
        _barrier(self.group)
        if self.rank == 0:
            write_manifest(directory, "tokstore", self.n_total, self.world)
        _barrier(self.group)

    @classmethod
    def load(cls, directory: str, device: int = 0, group=None,
             make_local: Optional[Callable] = None) -> "ShardedTokStore":
        """Token shards re-partitioned by doc id for the current world size (see ShardedIndex.load)."""
        from . import _lib

        man = read_manifest(directory, "tokstore")
        rank, world = _rank_world(group)
        lo, hi = shard_range(man["n_total"], rank, world)
        info = _lib.file_probe(os.path.join(directory, man["shards"][0]["file"]))
        if make_local is None:
            local = _lib.TokStore(info["dim"], _lib.DTYPE_NAMES[info["dtype"]], device, reserve_docs=hi - lo)
        else:
            local = make_local(info, hi - lo)
        for fname, first, n in plan_reshard(man["shards"], lo, hi):
            local.append_file(os.path.join(directory, fname), first, n)
        return cls(local, man["n_total"], group=group)

    def _p2p_setup(self, n_floats: int, device: torch.device):
        import torch.distributed._symmetric_memory as symm

        slot = (n_floats * 4 + 15) // 16 * 16
        flags_off = 2 * self.world * slot
        total = (flags_off + 2 * self.world * 4 + 15) // 16 * 16
        buf = symm.empty(total, dtype=torch.uint8, device=device)
        buf.zero_()
        hdl = symm.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
        bases = torch.tensor([int(x) for x in hdl.buffer_ptrs], dtype=torch.int64, device=device)
        torch.cuda.current_stream(device).synchronize()
        hdl.barrier()
        return {"key": n_floats, "buf": buf, "hdl": hdl, "bases": bases, "slot": slot, "flags_off": flags_off}

    def _maxsim_p2p(self, out: torch.Tensor) -> torch.Tensor:
        """Sum of the per-rank score matrices over peer memory: push + wait-sum kernels, no collective call."""
        import ctypes as C

        from . import _lib

        n = out.numel()
        st = getattr(self, "_p2p_state", None)
        if st is None or st["key"] != n:
            st = self._p2p_state = _agree_on_setup(lambda: self._p2p_setup(n, out.device), out.device, self.group)
            self._step = 0
        dev = out.device.index
        parity, seq = self._step & 1, (self._step % 0x7FFFFFFF) + 1
        self._step += 1
        stream = _lib._stream_ptr(dev)
        if st["slot"] != n * 4:                       # the push copies whole 16-byte units: give it a padded source
            src = torch.zeros(st["slot"] // 4, dtype=torch.float32, device=out.device)
            src[:n] = out.reshape(-1)
        else:
            src = out.contiguous()
        _lib.check(_lib.lib().ts_exchange_push(dev, C.c_void_p(src.data_ptr()), st["slot"], C.c_void_p(st["bases"].data_ptr()),
                                               self.world, self.rank, st["slot"], st["flags_off"], parity, seq, stream))
        res = torch.empty_like(out)
        _lib.check(_lib.lib().ts_exchange_wait_sum(dev, C.c_void_p(st["buf"].data_ptr()), self.world, n, st["slot"],
                                                   st["flags_off"], parity, seq, C.c_void_p(res.data_ptr()), stream))
        return res

    # ---- Stage 2 with the exchange fused into the scoring kernel (ts_maxsim_scatter) ---------------------------
    def _scatter_ok(self, q_tok: torch.Tensor) -> bool:
        """The same answer on every rank (it depends on shapes, the store's layout and the environment only)."""
        if self.world < 2 or not q_tok.is_cuda or getattr(self, "_p2p_off", False):
            return False
        env = os.environ.get("TS_P2P", "")
        if env == "0" or (env == "" and not P2P_DEFAULT):
            return False
        loc = self.local
        return (hasattr(loc, "maxsim_scatter") and getattr(loc, "layout", 0) == 1 and q_tok.dim() == 3
                and q_tok.shape[1] <= 128 and os.environ.get("TS_S2_SCATTER", "1") != "0")

    def _scatter_setup(self, n_floats: int, device: torch.device):
        """Symmetric receive buffer: two [B*C] fp32 matrices (step parity) + 2 x world u32 flags, zeroed."""
        import torch.distributed._symmetric_memory as symm

        slot = (n_floats * 4 + 15) // 16 * 16
        flags_off = 2 * slot
        total = (flags_off + 2 * self.world * 4 + 15) // 16 * 16
        buf = symm.empty(total, dtype=torch.uint8, device=device)
        buf.zero_()
        hdl = symm.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
        bases = torch.tensor([int(x) for x in hdl.buffer_ptrs], dtype=torch.int64, device=device)
        torch.cuda.current_stream(device).synchronize()
        hdl.barrier()
        return {"key": n_floats, "buf": buf, "hdl": hdl, "bases": bases, "slot": slot, "flags_off": flags_off}

    def _maxsim_scatter(self, q_tok: torch.Tensor, cand: torch.Tensor, **kw) -> torch.Tensor:
        """Every rank's kernel stores the scores of the candidates it owns into ALL ranks' matrices over NVLink and the
        last CTA publishes the step; the consumer kernel waits for the world's flags and takes the matrix.  Two
        launches of ours per step (+ the query prep), no collective call, nothing to reduce."""
        import ctypes as C

        from . import _lib

        B, Cn = cand.shape
        n = B * Cn
        st = getattr(self, "_sc_state", None)
        if st is None or st["key"] != n:
            st = self._sc_state = _agree_on_setup(lambda: self._scatter_setup(n, q_tok.device), q_tok.device, self.group)
            self._sc_step = 0
        parity, seq = self._sc_step & 1, (self._sc_step % 0x7FFFFFFF) + 1
        self._sc_step += 1
        mat_off, flags_off = parity * st["slot"], st["flags_off"] + parity * self.world * 4
        self.local.maxsim_scatter(q_tok, cand, st["bases"], self.world, self.rank, mat_off, flags_off, seq, **kw)
        res = torch.empty((B, Cn), dtype=torch.float32, device=q_tok.device)
        base = st["buf"].data_ptr()
        dev = q_tok.device.index
        _lib.check(_lib.lib().ts_exchange_wait_take(dev, C.c_void_p(base + mat_off), C.c_void_p(base + flags_off), self.world, seq,
                                                    n, C.c_void_p(res.data_ptr()), _lib._stream_ptr(dev)))
        return res

    def maxsim(self, q_tok: torch.Tensor, cand: torch.Tensor, **kw) -> torch.Tensor:
        if self._scatter_ok(q_tok):
            try:
                return self._maxsim_scatter(q_tok, cand, **kw)
            except PeerExchangeUnavailable as e:      # decided by all ranks together
                import logging

                logging.getLogger(__name__).warning(f"peer-memory exchange unavailable ({e}); using all-reduce")
                self._p2p_off = True
        out = self.local.maxsim(q_tok, cand, **kw)        # 0.0 for ids this shard does not own
        if self.world > 1:
            if out.is_cuda and os.environ.get("TS_P2P", "0") not in ("", "0") and not getattr(self, "_p2p_off", False):
                try:
                    return self._maxsim_p2p(out)
                except PeerExchangeUnavailable as e:
                    import logging

                    logging.getLogger(__name__).warning(f"peer-memory exchange unavailable ({e}); using all-reduce")
                    self._p2p_off = True
            dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.group)
        return out
