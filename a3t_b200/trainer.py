"""Data-parallel training step for the A3T model: one process per GPU, ONE all-reduce of one flat gradient buffer
(optionally issued range by range while the backward sweep is still running, `bucket_bytes > 0`).

Replaces the reference's trainer glue for this path (espnet2/train/trainer.py:243-275 DDP wrap,
:583-597 loss weighting, :631-675 clip / Adam / Noam):
  * every parameter is a view into one flat fp32 buffer; gradients into a second flat buffer
    whose 4-float tail carries the statistics the reference all-reduces separately
    (sum loss*B, sum loss_mlm*B, sum B, stop flag) -> a single in-place all-reduce per step (default).
    With `bucket_bytes > 0` the same buffer is cut into contiguous ranges of at least that size that are handed to
    NCCL (async, high-priority communicator of its own) as soon as the backward sweep has written them: the
    parameter order of the flat buffer is the reverse of the order the sweep finishes gradients in, so the finished
    region is always a suffix.  Measured on 8 x B200 (profiles/r02_scaling_overlap.md): NCCL's all-reduce kernels
    overlap the sweep but take 16-24 SMs, which costs the co-running tcgen05 GEMMs as much as the overlap hides
    (a 144-tile GEMM needs two waves on 124 SMs); with max_ctas = 4 nothing is disturbed but the last range is
    then exposed for 0.5 ms.  Net gain: +1 % at 2 GPUs, none at 8 -- hence opt-in;
  * gradient = sum_r(loss_r * B_r) / sum_r B_r, as trainer.py:583-595 + DDP mean produce;
  * clip_grad_norm_(max_norm) + Adam + NoamLR fused in one kernel pass (`a3t_adam_step`), with the
    non-finite-norm skip of trainer.py:640-656 decided on the device (no host sync);
  * BatchNorm running statistics stay rank-local (the reference has no SyncBN).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import _lib, graph


class DataParallelTrainer:
    def __init__(self, model, lr: float = 1.0, warmup: float = 4000.0, model_size: Optional[float] = None,
                 betas=(0.9, 0.999), eps: float = 1e-8, max_norm: float = 1.0, process_group=None, ops=None,
                 update_fn=None, bucket_bytes: int = 0, exchange_max_ctas: int = 0, last_range_full_speed: bool = False):
        """`ops` / `update_fn` exist for the CPU multi-process tests only (tests/test_dist_cpu.py passes the
        oracle backend and a torch restatement of `a3t_adam_step` to exercise the flat-buffer exchange
        over gloo); the product path leaves them None and requires a CUDA model."""
        self.model = model
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.xpg = process_group  # group the gradient ranges are exchanged on
        self.last_range_full_speed = bool(last_range_full_speed)
        self.lr, self.warmup, self.betas, self.eps, self.max_norm = lr, warmup, betas, eps, max_norm
        self.model_size = float(model_size if model_size is not None else model.encoder.attention_dim)
        params = [(n, p) for n, p in model.named_parameters()]
        dev = params[0][1].device
        if dev.type != "cuda" and (ops is None or update_fn is None):
            raise _lib.A3TError("DataParallelTrainer needs the model on a CUDA device")
        self.device = dev
        self.names = [n for n, _ in params]
        sizes = [p.numel() for _, p in params]
        self.n = sum(sizes)
        self.flat_p = torch.empty(self.n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(self.n + 4, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.gviews: Dict[str, torch.Tensor] = {}
        off = 0
        for (n, p), sz in zip(params, sizes):
            self.flat_p[off:off + sz].copy_(p.detach().reshape(-1))
            p.data = self.flat_p[off:off + sz].view(p.shape)
            self.gviews[n] = self.flat_g[off:off + sz].view(p.shape)
            off += sz
        if self.world > 1:  # C2: one parameter broadcast at init (trainer.py:250-265)
            dist.broadcast(self.flat_p, 0, group=self.pg)
            if bucket_bytes > 0 and dev.type == "cuda" and dist.get_backend(self.pg) == "nccl":
                # the range all-reduces run beside the backward kernels: on a HIGH-priority stream their CTAs are
                # placed as soon as any SM frees up (the default NCCL stream has normal priority and would queue
                # behind every already-launched compute kernel -- measured: no overlap at all at 8 GPUs)
                opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
                if exchange_max_ctas:
                    opts.config.max_ctas = int(exchange_max_ctas)
                ranks = dist.get_process_group_ranks(self.pg) if self.pg is not None else list(range(self.world))
                self.xpg = dist.new_group(ranks=ranks, backend="nccl", pg_options=opts)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)
        self.sq = torch.zeros(1, dtype=torch.float64, device=dev)
        self.sq_partial = torch.zeros(1024, dtype=torch.float64, device=dev)
        self.ops = ops if ops is not None else model._backend(dev)
        self._update_fn = update_fn
        self.stats = self.flat_g[self.n:]
        self._plan, self._plan_gen, self._qkv4 = None, -1, []
        self.in_place_repack = True  # False: drop the packed-weight cache after every step (lazy per-weight re-pack)
        self._rest = None
        self.bucket_floats = int(bucket_bytes) // 4  # 0 = one all-reduce after the sweep
        self._offsets = [0]
        for sz in sizes:
            self._offsets.append(self._offsets[-1] + sz)
        self.exchange_ranges = []  # [(lo, hi)] of the last step's all-reduce calls, in issue order

    def _exchange_hook(self, works):
        """Callback for `graph.backward(on_ready=...)`: copies the finished small gradients into the flat buffer
        and all-reduces (async, on NCCL's own stream) the finished suffix once it is at least one bucket long."""
        seen, state = set(), {"front": len(self.names), "hi": self.n + 4}
        self.exchange_ranges = []

        def hook(G, final=False):
            new = [n for n in G if n not in seen]
            cp = [n for n in new if G[n].data_ptr() != self.gviews[n].data_ptr()]
            if cp:
                torch._foreach_copy_([self.gviews[n] for n in cp], [G[n].view(self.gviews[n].shape) for n in cp])
            seen.update(new)
            i = state["front"]
            while i > 0 and self.names[i - 1] in seen:
                i -= 1
            lo, hi = (0 if final else self._offsets[i]), state["hi"]
            if final and i != 0:
                missing = [n for n in self.names[:i] if n not in seen]
                raise _lib.A3TError(f"backward produced no gradient for {missing[:4]}")
            if hi > lo and (final or hi - lo >= self.bucket_floats):
                # ranges that run beside the sweep go to the (CTA-limited, high-priority) exchange communicator; the
                # last one has nothing left to overlap with and takes the full-speed default communicator
                grp = self.pg if (final and self.last_range_full_speed) else self.xpg
                works.append(dist.all_reduce(self.flat_g[lo:hi], op=dist.ReduceOp.SUM, group=grp, async_op=True))
                self.exchange_ranges.append((lo, hi))
                state["front"], state["hi"] = i, lo

        return hook

    def step(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        """One optimizer step on this rank's shard of the global batch.  Returns the device tensor
        [sum loss*B, sum loss_mlm*B, sum B, stop] (all-reduced); no host synchronisation."""
        m = self.model
        ops, cfg, wc = self.ops, m.cfg, m._wcache
        P = m._param_dict()
        B = batch["speech"].shape[0]
        loss, before, after, ctx = graph.forward(ops, P, wc, cfg, batch, True, True)
        gloss = torch.full((1,), float(B), dtype=torch.float32, device=self.device)
        # one clear of the flat gradient buffer: weight-gradient GEMMs then write (split-K: accumulate) straight
        # into its views; only the small vectors (biases, norms, embeddings) are copied in afterwards
        self.flat_g.zero_()
        self.stats[0:1].copy_(loss).mul_(float(B))
        self.stats[1:2].copy_(loss).mul_(float(B))
        self.stats[2:3].fill_(float(B))
        self.stats[3:4].fill_(0.0)
        if self.world > 1 and self.bucket_floats > 0:
            # C3 (+C5/C6 piggy-backed), overlapped variant: the flat buffer (statistics tail included, with the first
            # range) is all-reduced range by range while the sweep runs; the step waits for all before the norm
            works = []
            hook = self._exchange_hook(works)
            G = graph.backward(ops, P, wc, cfg, ctx, gloss, gout=self.gviews, on_ready=hook)
            hook(G, final=True)
            for w in works:
                w.wait()
        else:
            G = graph.backward(ops, P, wc, cfg, ctx, gloss, gout=self.gviews)
            if self._rest is None:  # which gradients were not written in place is a property of the graph
                self._rest = [n for n in self.names if G[n].data_ptr() != self.gviews[n].data_ptr()]
            rest = self._rest
            torch._foreach_copy_([self.gviews[n] for n in rest], [G[n].view(self.gviews[n].shape) for n in rest])
            if self.world > 1:  # C3 (+C5/C6 piggy-backed): the single collective of the step
                dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.pg)
                self.exchange_ranges = [(0, self.n + 4)]
        if self._update_fn is not None:  # CPU test seam
            self._update_fn(self)
            wc.clear()
            return self.stats
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.call("a3t_grad_sqnorm", self.flat_g.data_ptr(), self.n, self.sq.data_ptr(), self.sq_partial.data_ptr(), st)
        _lib.call("a3t_adam_step", self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.flat_m.data_ptr(),
                  self.flat_v.data_ptr(), self.n, self.sq.data_ptr(), self.step_count.data_ptr(), self.lr,
                  self.model_size, self.warmup, self.betas[0], self.betas[1], self.eps, self.max_norm, 1.0,
                  self.stats[2:3].data_ptr(), st)
        self._refresh_packed_weights(P, wc)  # parameters changed under the packed bf16 copies
        ops.advance_seed()
        return self.stats

    def _refresh_packed_weights(self, P, wc):
        """The bf16 GEMM operand copies of every weight are rewritten IN PLACE by one batched kernel (the
        parameters live at fixed addresses inside `flat_p`), so the weight cache stays valid and the next
        step launches no per-weight packing kernels.  Falls back to dropping the cache (lazy re-pack).

        The plan holds raw pointers to the cache's bf16 buffers, so it is rebuilt whenever the cache rebuilt an
        entry since the plan was made (`load_state_dict` / checkpoint resume / any in-place torch write bumps the
        parameter version -> cache miss -> new buffers -> `wc.generation` changes)."""
        ops = self.ops
        if not self.in_place_repack:
            self._plan = None
            wc.clear()
            return
        if self._plan_gen != wc.generation:
            self._plan_gen = wc.generation
            self._plan, self._qkv4 = None, []
            entries, qkv4 = [], []
            for key, (_, val) in wc._c.items():
                if key.endswith(".qkv4"):
                    a = key[:-len(".qkv4")]
                    pw, b4 = val
                    wq, wk, wv = (P[f"{a}.linear_{n}.weight"] for n in "qkv")
                    entries.append(([wq, wq, wk, wv], pw))
                    qkv4.append((P[f"{a}.linear_q.bias"], P[f"{a}.linear_k.bias"], P[f"{a}.linear_v.bias"],
                                 P[f"{a}.pos_bias_u"], P[f"{a}.pos_bias_v"], b4))
                else:
                    entries.append(([val.w], val))
            plan = ops.build_pack_plan(entries) if hasattr(ops, "build_pack_plan") else None
            if plan is not None:
                self._plan, self._qkv4 = plan, qkv4
        if self._plan is None:
            wc.clear()
            return
        ops.repack(self._plan)
        for bq, bk, bv, u, v, b4 in self._qkv4:
            ops.qkv4_bias(bq, bk, bv, u, v, b4)
