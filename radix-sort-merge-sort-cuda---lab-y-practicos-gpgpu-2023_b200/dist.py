"""One-box multi-GPU sort: MSD partition + exchange over NVLink + local radix sort.

One process per GPU (``torchrun``); ``torch.distributed`` is the plumbing (NCCL on GPUs, gloo in
the CPU tests).  No reference counterpart: the lab is single-GPU (SRM/run.sh:11); this is
north_star (c) / SURVEY.md section 8(e).

    phase 1  every rank histograms the top ``bits`` bits of its keys      (CUDA, b200sort_dist_histogram_i32)
    phase 2  counts are all-gathered over NCCL (their sum is the all-reduce north_star names); the planner
             gives each rank one contiguous value range of ~total/world keys.  It runs ON THE DEVICE
             (b200sort_dist_plan_device, one CTA) and leaves offsets and counts in a device record, so the
             whole sort is enqueued without the host ever reading a count; the pure host planner
             (b200sort_dist_plan, same boundaries bit for bit) serves the NCCL baseline and the CPU tests
    phase 3  every rank multisplits its keys by destination              (CUDA, b200sort_dist_partition_i32)
               exchange="p2p"   the destination table holds the peers' receive buffers (CUDA IPC
                                mapped): the kernel's coalesced stores ARE the exchange and cross
                                NVLink while partitioning continues  -- the fused path, default;
               exchange="nccl"  the kernel fills a local send buffer and an NCCL all-to-all moves
                                the blocks -- the baseline the fused path is measured against;
    phase 4  every rank radix-sorts what it received, the key count read from the device record
                                                                          (CUDA, b200sort_radix_copy_devn_i32)

Rank r ends up holding the r-th contiguous slice of the global order (sizes differ by the
granularity of the 2^bits bins); concatenating the ranks' outputs gives the sorted array.
"""
from __future__ import annotations

import ctypes
import json
import os
import time

import numpy as np

from ._lib import ALGO_RADIX, check, lib

DEFAULT_BITS = 14          # 2^14 bins: a byte that holds 90 % of the keys still splits within 6 % of a rank's share
PLAN_BYTES = 1424          # B200SORT_DIST_PLAN_BYTES
PLAN_TOP_HIST_OFFSET = 400 # uint32 top_hist[256]
PLAN_M_OFFSET = 384        # uint32 m, then uint32 error


# ---- host-side pieces (also exercised on CPU by the gloo tests) ------------------------------------

def plan(all_hist: np.ndarray, rank: int, bits: int):
    """b200sort_dist_plan: (bin_owner int32[nbins], recv_count, send_count, dst_offset uint64[world])."""
    all_hist = np.ascontiguousarray(all_hist, dtype=np.uint64)
    world, nbins = all_hist.shape
    assert nbins == 1 << bits
    owner = np.zeros(nbins, dtype=np.int32)
    recv = np.zeros(world, dtype=np.uint64)
    send = np.zeros(world, dtype=np.uint64)
    offs = np.zeros(world, dtype=np.uint64)
    check(lib().b200sort_dist_plan(all_hist.ctypes.data, world, rank, bits, owner.ctypes.data,
                                   recv.ctypes.data, send.ctypes.data, offs.ctypes.data))
    return owner, recv, send, offs


def host_histogram(keys: np.ndarray, bits: int) -> np.ndarray:
    """numpy twin of b200sort_dist_histogram_i32 (CPU tests of the host logic only)."""
    top = (keys.view(np.uint32) ^ np.uint32(0x80000000)) >> np.uint32(32 - bits)
    return np.bincount(top.astype(np.int64), minlength=1 << bits).astype(np.uint64)


# ---- the GPU path ---------------------------------------------------------------------------------------

class DistSorter:
    """Owns this rank's receive / scratch / output buffers and the peer mappings."""

    def __init__(self, n_local: int, bits: int = DEFAULT_BITS, exchange: str = "p2p", headroom: float = 1.25,
                 group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.bits, self.nbins = bits, 1 << bits
        self.exchange = exchange
        self.n_local = int(n_local)
        self.cap = int(self.n_local * headroom) + 4096
        self.dev = torch.device("cuda", torch.cuda.current_device())
        L = lib()
        check(L.b200sort_device_check())
        # receive buffer: a whole cudaMalloc allocation so that it can be exported over CUDA IPC
        p = ctypes.c_void_p()
        check(L.b200sort_device_malloc(ctypes.byref(p), self.cap * 4))
        self.recv_ptr = p.value
        self.tmp = torch.empty(self.cap, dtype=torch.int32, device=self.dev)
        self.out = torch.empty(self.cap, dtype=torch.int32, device=self.dev)
        self.hist = torch.zeros(self.nbins, dtype=torch.int64, device=self.dev)
        self.all_hist = torch.zeros(self.world * self.nbins, dtype=torch.int64, device=self.dev)
        self.owner_dev = torch.zeros(self.nbins, dtype=torch.int32, device=self.dev)
        ws = max(L.b200sort_workspace_bytes(self.cap, ALGO_RADIX), L.b200sort_dist_workspace_bytes(self.cap, bits))
        self.ws = torch.empty(ws + 512, dtype=torch.uint8, device=self.dev)
        self.ws_ptr = self.ws.data_ptr() + (-self.ws.data_ptr()) % 256
        self.ws_bytes = ws
        self.part_ws_bytes = L.b200sort_dist_workspace_bytes(self.cap, bits)
        self.part_ws = torch.zeros(self.part_ws_bytes + 256, dtype=torch.uint8, device=self.dev)
        self.part_ws_ptr = self.part_ws.data_ptr() + (-self.part_ws.data_ptr()) % 256
        self.plan_dev = torch.zeros(PLAN_BYTES // 8, dtype=torch.int64, device=self.dev)     # the device plan record
        # digit histograms counted at the source: [destination][4][256], and the row that is mine after the reduce-scatter
        self.src_hist = torch.zeros(self.world * 1024, dtype=torch.int32, device=self.dev)
        self.my_hist = torch.zeros(1024, dtype=torch.int32, device=self.dev)
        # Worth it where the exchange is bound by the fabric, not by the partition kernel (4+ GPUs): the four shared-memory
        # atomics per key cost the kernel 0.4 ms at 2^28 keys, which 2 GPUs do not hide (measured: 4.43 ms either way).
        env = os.environ.get("B200SORT_DIST_HIST_AT_SOURCE")
        self.hist_at_source = ((env != "0") if env is not None else (self.world >= 4)) and bits >= 8 and exchange == "p2p"
        self.flag = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.peer_ptrs = [None] * self.world
        self.send = None
        if exchange == "p2p":
            handle = (ctypes.c_ubyte * 64)()
            check(L.b200sort_ipc_export(self.recv_ptr, handle))
            mine = torch.tensor(list(handle), dtype=torch.uint8, device=self.dev)
            everyone = [torch.zeros(64, dtype=torch.uint8, device=self.dev) for _ in range(self.world)]
            dist.all_gather(everyone, mine, group=group)
            for r in range(self.world):
                if r == self.rank:
                    self.peer_ptrs[r] = self.recv_ptr
                else:
                    h = (ctypes.c_ubyte * 64)(*everyone[r].cpu().tolist())
                    q = ctypes.c_void_p()
                    check(L.b200sort_ipc_open(h, ctypes.byref(q)))
                    self.peer_ptrs[r] = q.value
        elif exchange == "nccl":
            self.send = torch.empty(self.n_local, dtype=torch.int32, device=self.dev)
            self.recv_t = torch.empty(self.cap, dtype=torch.int32, device=self.dev)
        else:
            raise ValueError("exchange must be 'p2p' or 'nccl'")
        self.last = {}

    def close(self) -> None:
        L = lib()
        self.torch.cuda.synchronize()
        self.dist.barrier(group=self.group)
        for r, p in enumerate(self.peer_ptrs):
            if p is not None and r != self.rank:
                L.b200sort_ipc_close(p)
        self.peer_ptrs = [None] * self.world
        self.dist.barrier(group=self.group)
        if self.recv_ptr:
            L.b200sort_device_free(self.recv_ptr)
            self.recv_ptr = None

    def sort(self, keys, phase_events=None):
        """Sort the distributed array whose local part is the int32 CUDA tensor ``keys`` (read only).
        Returns (this rank's slice of the global order as a tensor view, its length).  ``sync=False`` via
        ``sort_async`` returns the full output buffer and the device record instead, without any host
        synchronisation.  ``phase_events``: optional list that receives CUDA events recorded at the phase
        boundaries (start, histogram, all-gather + plan, partition + exchange, local sort)."""
        out = self.sort_async(keys, phase_events)
        m = self.count()
        return out[:m], m

    def count(self) -> int:
        """Keys this rank owns after the last sort (synchronises; raises if a receive buffer was too small)."""
        rec = self.plan_dev.cpu().numpy()
        if self.exchange == "nccl":
            return self.last["recv"]
        words = rec.view(np.uint32)
        m, err = int(words[PLAN_M_OFFSET // 4]), int(words[PLAN_M_OFFSET // 4 + 1])
        recv = rec.view(np.uint64)[:self.world]
        if err:
            raise RuntimeError(f"a rank would receive {int(recv.max())} keys, receive buffers hold {self.cap}: "
                               "raise headroom (skewed keys)")
        send = rec.view(np.uint64)[16:16 + self.world]
        self.last = {"recv": m, "sent_remote": int(send.sum() - send[self.rank]), "recv_counts": recv.copy()}
        return m

    def sort_async(self, keys, phase_events=None):
        torch, dist, L = self.torch, self.dist, lib()

        def mark():
            if phase_events is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                phase_events.append(e)
        mark()
        n = keys.numel()
        assert n <= self.n_local and keys.dtype == torch.int32 and keys.is_cuda and keys.is_contiguous()
        stream = torch.cuda.current_stream().cuda_stream
        # phase 1
        check(L.b200sort_dist_histogram_i32(keys.data_ptr(), n, self.bits, self.hist.data_ptr(), stream))
        mark()
        # phase 2: counts of every rank (world x nbins, 8 B each: latency-bound).  The all-gather is also the
        # barrier that protects the receive buffers: it completes on a rank only when every rank has enqueued it,
        # i.e. after every rank's previous local sort (earlier in its stream) has read what it received, and this
        # sort's peer writes come after it in stream order.
        dist.all_gather_into_tensor(self.all_hist, self.hist, group=self.group)
        if self.exchange == "p2p":
            check(L.b200sort_dist_plan_device(self.all_hist.data_ptr(), self.world, self.rank, self.bits, self.cap,
                                              self.owner_dev.data_ptr(), self.plan_dev.data_ptr(), self.part_ws_ptr,
                                              self.part_ws_bytes, stream))
            mark()
            # phase 3: the kernel's bulk copies into the peers' receive buffers ARE the exchange
            base = (ctypes.c_void_p * self.world)(*self.peer_ptrs)
            hist = self.hist_at_source
            check(L.b200sort_dist_partition_planned_i32(keys.data_ptr(), n, self.bits, self.world, base,
                                                        self.owner_dev.data_ptr(), self.plan_dev.data_ptr(),
                                                        self.src_hist.data_ptr() if hist else None,
                                                        self.part_ws_ptr, self.part_ws_bytes, stream))
            # stream-ordered barrier: nobody sorts before every peer's stores have landed.  With the digit histograms
            # counted at the source the barrier carries them: a reduce-scatter hands every rank the histogram of
            # exactly the keys it received, and its local sort skips its own histogram kernel.
            if hist:
                dist.reduce_scatter_tensor(self.my_hist, self.src_hist, group=self.group)
                # digits 0..2 were counted at the source; the top byte's histogram follows from the bin counts (plan)
                self.my_hist[768:].copy_(self.plan_dev.view(torch.int32)[PLAN_TOP_HIST_OFFSET // 4:PLAN_TOP_HIST_OFFSET // 4 + 256])
            else:
                dist.all_reduce(self.flag, group=self.group)
            mark()
            # phase 4: the key count comes from the device record
            check(L.b200sort_radix_copy_devn_i32(self.recv_ptr, self.out.data_ptr(), self.tmp.data_ptr(), self.cap,
                                                 self.plan_dev.data_ptr() + PLAN_M_OFFSET,
                                                 self.my_hist.data_ptr() if hist else None,
                                                 self.ws_ptr, self.ws_bytes, stream))
            mark()
            return self.out
        # ---- the NCCL all-to-all baseline: host planner (the collective needs the split sizes on the host) ----
        all_hist = self.all_hist.cpu().numpy().astype(np.uint64).reshape(self.world, self.nbins)
        owner, recv, send, offs = plan(all_hist, self.rank, self.bits)
        m = int(recv[self.rank])
        if int(recv.max()) > self.cap:
            raise RuntimeError(f"rank would receive {int(recv.max())} keys, receive buffers hold {self.cap}: "
                               "raise headroom (skewed keys)")
        self.owner_dev.copy_(torch.from_numpy(owner), non_blocking=False)
        mark()
        send_disp = np.concatenate([[0], np.cumsum(send)[:-1]]).astype(np.uint64)
        base = (ctypes.c_void_p * self.world)(*([self.send.data_ptr()] * self.world))
        check(L.b200sort_dist_partition_i32(keys.data_ptr(), n, self.bits, self.world, base,
                                            self.owner_dev.data_ptr(), send_disp.ctypes.data,
                                            self.part_ws_ptr, self.part_ws_bytes, stream))
        in_splits = [int(all_hist[s][owner == self.rank].sum()) for s in range(self.world)]
        out_splits = [int(x) for x in send]
        dist.all_to_all_single(self.recv_t[:m], self.send[:n], in_splits, out_splits, group=self.group)
        mark()
        check(L.b200sort_sort_copy_i32(ALGO_RADIX, self.recv_t.data_ptr(), self.out.data_ptr(), self.tmp.data_ptr(), m,
                                       self.ws_ptr, self.ws_bytes, stream))
        mark()
        self.last = {"recv": m, "sent_remote": int(send.sum() - send[self.rank]), "recv_counts": recv}
        return self.out


# ---- bench.py --gpus N -----------------------------------------------------------------------------------

def _make_keys(torch, dev, n_local: int, dist_name: str, seed: int):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    src = torch.randint(-2**31, 2**31, (n_local,), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
    if dist_name == "uniform":
        return src
    if dist_name == "skewed90":
        hot = torch.rand(n_local, device=dev, generator=g) < 0.9
        return torch.where(hot, (src & 0x00FFFFFF) | 0x40000000, src)
    raise SystemExit("multi-GPU bench supports --dist uniform|skewed90")


def _global_check(torch, dist, dev, world, src, out, m):
    """Per-rank sortedness, ordered rank boundaries, key count and multiset sum over all ranks."""
    ok = bool((out[1:] >= out[:-1]).all().item()) if m > 1 else True
    lo = out[0].item() if m > 0 else 2**31 - 1
    hi = out[m - 1].item() if m > 0 else -2**31
    edges = torch.tensor([lo, hi, m, int(out.sum(dtype=torch.int64).item()) if m else 0], dtype=torch.int64, device=dev)
    allv = [torch.zeros_like(edges) for _ in range(world)]
    dist.all_gather(allv, edges)
    tot_in = torch.tensor([int(src.sum(dtype=torch.int64).item())], dtype=torch.int64, device=dev)
    dist.all_reduce(tot_in)
    prev_hi = -2**31
    for e in allv:
        if int(e[2]) > 0:
            assert int(e[0]) >= prev_hi, "bench: rank boundaries out of order"
            prev_hi = int(e[1])
    assert ok, "bench: a rank's slice is not sorted"
    assert sum(int(e[2]) for e in allv) == world * src.numel(), "bench: key count changed"
    assert sum(int(e[3]) for e in allv) == int(tot_in.item()), "bench: multiset sum changed"
    return [int(e[2]) for e in allv]


def _timed_sorts(torch, dist, dev, sorter, src, steps: int, warmup: int):
    """`steps` distributed sorts between two CUDA events (nothing synchronises with the host inside), max over
    ranks; then the phase breakdown of one more sort.  Returns (ms_per_step, phases, recv_counts, sent_remote)."""
    world = sorter.world
    for _ in range(max(warmup, 3)):
        sorter.sort_async(src)
    torch.cuda.synchronize()
    m = sorter.count()
    recv_counts = _global_check(torch, dist, dev, world, src, sorter.out[:m], m)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        sorter.sort_async(src)
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    m = sorter.count()
    _global_check(torch, dist, dev, world, src, sorter.out[:m], m)
    evs = []
    sorter.sort_async(src, phase_events=evs)
    torch.cuda.synchronize()
    ph = torch.tensor([evs[i].elapsed_time(evs[i + 1]) for i in range(4)], dtype=torch.float64, device=dev)
    dist.all_reduce(ph, op=dist.ReduceOp.MAX)
    phases = {k: float(v) for k, v in zip(("msd_histogram_ms", "allgather_plan_ms", "partition_exchange_ms", "local_sort_ms"), ph.tolist())}
    return float(ms.item()) / steps, phases, recv_counts, sorter.last.get("sent_remote", 0)


def bench_main(args, metric: str, unit: str, ClockSampler, peaks_fn) -> None:
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if rank == 0:
            print(json.dumps({"error": f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})"}))
        return
    torch.cuda.set_device(local)
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"          # keep rank 0's stdout to the one JSON line
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    L = lib()
    check(L.b200sort_device_check())               # per-device initialisation (self-test) outside the timed region
    # --total-log2n T: strong scaling, 2^T keys in total (BASELINE config 5: T = 30); default: 2^log2n per GPU
    if args.total_log2n is not None:
        n_local = (1 << args.total_log2n) // world
        scaling = "strong"
    else:
        n_local = 1 << args.log2n
        scaling = "weak"
    src = _make_keys(torch, dev, n_local, args.dist, 1000 + rank)
    sorter = DistSorter(n_local, exchange=args.exchange, headroom=1.25 if args.dist == "uniform" else 1.5)

    for _ in range(max(args.warmup, 3)):
        sorter.sort_async(src)
    torch.cuda.synchronize()
    m = sorter.count()
    _global_check(torch, dist, dev, world, src, sorter.out[:m], m)

    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.02)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    L.b200sort_launch_count_reset()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        sorter.sort_async(src)
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    t1 = time.perf_counter()
    launches = int(L.b200sort_launch_count())
    sampler.stop_flag = True
    sampler.join(timeout=2)
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    ms_per_step = ms_total / args.steps
    m = sorter.count()
    recv_all = _global_check(torch, dist, dev, world, src, sorter.out[:m], m)
    total_keys = world * n_local
    sent = sorter.last.get("sent_remote", 0)
    # phase breakdown of one more sort (CUDA events at the phase boundaries, max over ranks)
    evs = []
    sorter.sort_async(src, phase_events=evs)
    torch.cuda.synchronize()
    ph = torch.tensor([evs[i].elapsed_time(evs[i + 1]) for i in range(4)], dtype=torch.float64, device=dev)
    dist.all_reduce(ph, op=dist.ReduceOp.MAX)
    phases = {k: float(v) for k, v in zip(("msd_histogram_ms", "allgather_plan_ms", "partition_exchange_ms", "local_sort_ms"), ph.tolist())}

    # e2e: host buffers in, host buffers out (pinned), every step
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    h_in = torch.empty(n_local, dtype=torch.int32, pin_memory=True)
    h_in.copy_(src)
    h_out = torch.empty(sorter.cap, dtype=torch.int32, pin_memory=True)
    d_in = torch.empty_like(src)
    torch.cuda.synchronize()
    dist.barrier()
    el = 0.0
    for i in range(e2e_steps + 1):
        dist.barrier()
        t = time.perf_counter()
        d_in.copy_(h_in, non_blocking=True)
        o, mm = sorter.sort(d_in)
        h_out[:mm].copy_(o, non_blocking=True)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if i > 0:
            el += float(dt.item())
    e2e_value = total_keys * e2e_steps / el
    del h_in, h_out, d_in

    # ---- BASELINE config 5 rows, measured in every multi-GPU run: 2^30 keys IN TOTAL split over the ranks
    # (uniform and skewed90), and the skewed90 run at this run's own size ------------------------------------------
    config5 = []
    if not args.no_configs:
        cases = [("skewed90 at this run's size", n_local, "skewed90")]
        if (1 << 30) // world != n_local:
            cases.insert(0, ("config5: 2^30 uniform keys in total", (1 << 30) // world, "uniform"))
        else:
            cases.insert(0, ("config5: 2^30 uniform keys in total", None, None))       # the headline run IS that case
        if (1 << 30) // world != n_local:
            cases.append(("config5: 2^30 skewed90 keys in total", (1 << 30) // world, "skewed90"))
        cur_sorter, cur_n = sorter, n_local
        for name, nl, dname in cases:
            if nl is None:
                config5.append({"config": name, "n_per_gpu": n_local, "dist": "uniform", "ms_per_step": ms_per_step,
                                "keys_per_s": total_keys / (ms_per_step / 1e3), "phases_max_over_ranks": phases,
                                "recv_counts": recv_all, "note": "the headline run"})
                continue
            try:
                if nl != cur_n:
                    if cur_sorter is not sorter:
                        cur_sorter.close()
                    cur_sorter, cur_n = DistSorter(nl, exchange=args.exchange, headroom=1.5), nl
                elif dname != "uniform" and cur_sorter is sorter and sorter.cap < int(n_local * 1.4):
                    cur_sorter, cur_n = DistSorter(nl, exchange=args.exchange, headroom=1.5), nl
                keys = _make_keys(torch, dev, nl, dname, 2000 + rank)
                msx, phx, recvx, _ = _timed_sorts(torch, dist, dev, cur_sorter, keys, args.config_steps, 3)
                mean = sum(recvx) / len(recvx)
                config5.append({"config": name, "n_per_gpu": nl, "dist": dname, "ms_per_step": msx,
                                "keys_per_s": world * nl / (msx / 1e3), "phases_max_over_ranks": phx,
                                "recv_counts": recvx, "recv_max_over_mean": max(recvx) / mean if mean else None})
                del keys
            except Exception as e:           # a row that cannot run says so
                config5.append({"config": name, "error": repr(e)[:300]})
        if cur_sorter is not sorter:
            cur_sorter.close()

    if rank == 0:
        peaks = peaks_fn()
        line = {
            "metric": metric, "value": total_keys / (ms_per_step / 1e3), "unit": unit, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "int32",
            "data": "synthetic",
            "config": {"workload": f"distributed radix sort, n={n_local} {args.dist} int32 keys PER GPU "
                                   f"({world} x {n_local} = {total_keys} keys), MSD partition on the top "
                                   f"{sorter.bits} bits + exchange + local onesweep",
                       "exchange": "fused scatter over NVLink: TMA bulk copies into CUDA-IPC mapped receive buffers; "
                                   "device planner, no host synchronisation inside a sort"
                                   if args.exchange == "p2p" else "NCCL all_to_all_single (host planner)",
                       "dist": args.dist, "seed": "1000+rank", "n_per_gpu": n_local,
                       "l2": "inputs larger than L2" if 4 * n_local > 126e6 else "inputs fit L2", "recv_counts": recv_all,
                       "recv_max_over_mean": max(recv_all) / (sum(recv_all) / len(recv_all))},
            "roofline": {"bound": "hbm", "achieved": 48.0 * n_local / (ms_per_step / 1e3) / 1e9,
                         "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": 48.0 * n_local / (ms_per_step / 1e3) / 1e9 / peaks["hbm_gbs"], "traffic": None,
                         "kernel": "whole distributed sort per GPU: 4 (MSD histogram) + 8 (partition/exchange) + "
                                   "36 (local sort) = 48 B/key of HBM traffic",
                         "peak_source": peaks["source"],
                         "nvlink_bytes_sent_per_gpu": 4 * sent, "phases_max_over_ranks": phases,
                         "nvlink_gbs_per_gpu_out": 4 * sent / (phases["partition_exchange_ms"] / 1e3) / 1e9},
            "cpu_baseline": None,
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": 4 * n_local * world,
                    "d2h_bytes_per_step": 4 * n_local * world, "steps": e2e_steps,
                    "api": "DistSorter.sort on pinned host buffers: H2D + distributed sort + D2H per rank"},
            "gpu_launches": launches * world, "clocks": sampler.summary(t0, t1), "configs": config5,
        }
        print(json.dumps(line), flush=True)
    sorter.close()
    dist.destroy_process_group()
