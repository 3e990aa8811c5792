#!/usr/bin/env python
"""Benchmark of the mSWE-GNN rollout hot path (contract: see the task statement / DESIGN.md §5).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3]

One "step" = one autoregressive rollout step (boundary injection, full MSGNN forward, window
shift) of the default config.yaml mSWE-GNN (K=4, F=64, 3-layer MLPs, 4 scales) on a synthetic
triangular mesh.  N=1: cfg3 = tri(712,712), 1,346,574 nodes (BASELINE.json configs[2], the
single-GPU configuration the rollout metric is quoted on).  Metric: node-steps/s.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
    os.environ["NCCL_DEBUG"] = "NONE"                  # NCCL prints its version banner to stdout at these two levels;
                                                       # rank 0 must print ONE JSON line (INFO and above are left alone)

import torch  # noqa: E402

MODEL_CFG = dict(hid_features=64, mlp_layers=3, seed=666, learned_residuals=True, mlp_activation="prelu",
                 gnn_activation="tanh", edge_mlp=True, normalize=True, with_filter_matrix=True, with_gradient=True,
                 with_WL=True, K=4, learned_pooling=False, skip_connections=True)      # config.yaml:42-58
CTOR = dict(num_node_features=8, num_edge_features=1, num_scales=4, previous_t=3, **MODEL_CFG)
WORKLOADS = {"cfg1": (32, 24), "cfg5": (224, 224), "cfg3": (712, 712), "cfg4": (2832, 2832), "cpu_sample": (192, 192)}
F, K_HOPS, S = 64, 4, 4


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return dict(hbm=p["hbm_gbs"], hbm_src="measured", bf16=p["bf16_tflops_sustained"], bf16_burst=p["bf16_tflops"])
    except Exception:
        return dict(hbm=6650.0, hbm_src="fallback", bf16=1400.0, bf16_burst=1590.0)


def algorithmic_bytes(nx, ny):
    """per-forward algorithmic bytes (fp32 features, int32 CSR, neighbour gathers
    L2-served, s_ij written once per SWEGNN call and re-read every hop)."""
    from mswe_gnn_b200.utils.synthetic import tri_level_sizes
    sz = tri_level_sizes(nx, ny, S)
    N = [s[0] for s in sz]; E = [s[1] for s in sz]; I = [s[2] for s in sz]
    gate = lambda s: 4 * F * (2 * N[s] + 2 * E[s]) + 4 * (E[s] + N[s] + 1)
    hop = lambda s: 4 * F * (E[s] + 2 * N[s]) + 4 * (E[s] + N[s] + 1)
    calls = list(range(S - 1)) + list(range(S - 1, -1, -1))
    total = sum(gate(s) + K_HOPS * hop(s) for s in calls)
    Nt, Et = sum(N), sum(E)
    total += 4 * (Et * (1 + F) + Nt * (9 + 2 * F) + Nt * (F + 2))
    for s in range(S - 1):
        total += 4 * F * (2 * (N[s] + N[s + 1]) + 2 * N[s]) + 8 * I[s] + 4 * F * (N[s] + N[s + 1]) + 4 * I[s]
    return dict(total=total, gate=[gate(s) for s in range(S)], hop=[hop(s) for s in range(S)], N=N, E=E, I=I)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx = float(parts[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port of the reference's PyG path on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_reference_rate(steps, warmup, threads=None, wl="cfg3"):
    """node-steps/s of the reference CPU path (oracle/swe_oracle.py: per-hop masks, compaction and
    K× edge-MLP evaluation exactly as /root/reference/models/gnn.py does) on the headline mesh itself
    (cfg3 = tri(712,712): about half a minute per step on 16 host threads), bounded in the number of steps."""
    from oracle import swe_oracle as O
    from mswe_gnn_b200.models.gnn import MSGNN
    from mswe_gnn_b200.utils.synthetic import make_tri_mesh
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    nx, ny = WORKLOADS[wl]
    data = make_tri_mesh(nx, ny, S, rollout_steps=steps + warmup)
    sd = MSGNN(**CTOR).state_dict()
    spec = O.ModelSpec("MSGNN", **CTOR)
    n = data.x.shape[0]
    t = O._Bag(**{k: (getattr(data, k).clone() if torch.is_tensor(getattr(data, k)) else getattr(data, k)) for k in data.keys()})
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            t.x[:, -6:] = O.apply_boundary_condition(t.x[:, -6:], t.BC[:, :, i], t.node_BC, t.type_BC)
            p = O.forward(sd, spec, t)
            t.x = O.use_prediction(t.x, p, 3)
            times.append(time.perf_counter() - t0)
    timed = times[warmup:]
    sec = sum(timed) / len(timed)
    return dict(value=n / sec, ms_per_step=sec * 1e3, cores=threads, nodes=n,
                sample=f"{steps} rollout steps of the same model on tri({nx},{ny}) 4-scale ({n} nodes), torch {torch.__version__} CPU, {threads} threads")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload in TRAIN_WORKLOADS:
        r = cpu_training_rate(args.workload)
        line = {"impl": "reference", "metric": "mSWE-GNN training node-steps/sec (forward+backward+AdamW)", "value": r["value"],
                "unit": "node-steps/s", "n_gpus": args.gpus, "steps": 1, "warmup": 1, "ms_per_step": None, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{args.workload}: CPU arm timed on a bounded sample: {r['sample']}"},
                "cpu_baseline": {"value": r["value"], "unit": "node-steps/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": "node-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return
    steps, warmup = min(args.steps, 2), min(args.warmup, 1)
    # the CPU arm runs the headline mesh itself (cfg3); cfg4 (21 M nodes, ~10 min per step) is timed on cfg3 and says so
    r = cpu_reference_rate(steps, warmup, wl="cfg3" if (args.workload or "cfg3") in ("cfg3", "cfg4") else args.workload)
    nx, ny = WORKLOADS[args.workload or "cfg3"]
    line = {"impl": "reference", "metric": "mSWE-GNN rollout node-steps/sec", "value": r["value"], "unit": "node-steps/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload or 'cfg3'}: default config.yaml mSWE-GNN (K4,F64,S4) rollout on tri({nx},{ny}); "
                                   f"CPU arm timed on a bounded sample: {r['sample']}"},
            "cpu_baseline": {"value": r["value"], "unit": "node-steps/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "node-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def mesh_for(world, wl):
    """Workload mesh.  N=1: the named config.  N>1 (default workload): ONE partitioned mesh with a fixed share of
    2 x cfg3's nodes per GPU (weak scaling), tri(712a, 712b) with a*b = 2N — so that N=8 is tri(2848,2848) = 21.5 M
    nodes, the >= 16 M-node mesh of BASELINE.json configs[3] (cfg4 itself, tri(2832,2832), is `--workload cfg4`)."""
    nx, ny = WORKLOADS[wl]
    if world > 1 and wl == "cfg3":
        a = 1
        while a * a < 2 * world:
            a *= 2
        b = 2 * world // a
        nx, ny = nx * a, ny * b
    return nx, ny


def partition_parity_check(model, dev, world, rank, full_graph=None, steps=3):
    """Before anything is timed: a mesh partitioned over the ranks (peer-memory halo exchange, device-side flags, captured
    step) against the same rollout un-partitioned on one GPU, owned rows compared BIT FOR BIT on every rank.  Small mesh
    by default (every rank runs the un-partitioned rollout itself); with `full_graph` (cfg4) rank 0 runs the whole mesh
    and broadcasts its predictions."""
    import torch.distributed as dist
    from mswe_gnn_b200.parallel import PartitionedRollout
    from mswe_gnn_b200.training.train import rollout_test
    from mswe_gnn_b200.utils.synthetic import make_tri_mesh
    if full_graph is None:
        nx, ny = 160, 96
        g = make_tri_mesh(nx, ny, S, rollout_steps=steps, seed=3, with_y=False)
        g.y = torch.empty(0, 2, steps)
        name = f"tri({nx},{ny})"
    else:
        g, name = full_graph, "the benchmark mesh itself"
    pr = PartitionedRollout(model, g, steps, dev, transport="peer")
    preds = pr.run()
    mine, gids = pr.owned_predictions()
    mine = mine.clone()
    torch.cuda.synchronize()
    pr.close()
    if full_graph is None:
        ref = rollout_test(model, g.to(dev)).permute(2, 0, 1).contiguous()                  # [T, N, 2]
    else:
        n = int(g.x.shape[0])
        ref = torch.empty(steps, n, 2, device=dev)
        if rank == 0:
            gg = g.to(dev)
            gg.y = torch.empty(0, 2, steps)
            ref.copy_(rollout_test(model, gg).permute(2, 0, 1))
            del gg
        dist.broadcast(ref, 0)
    same = torch.equal(mine, ref[:, torch.from_numpy(gids).to(dev)])
    flag = torch.tensor([1 if same else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    del ref
    torch.cuda.empty_cache()
    return {"peer_bit_exact": bool(flag.item()), "mesh": name, "nodes": int(g.x.shape[0]), "steps": steps, "ranks": world,
            "what": "owned rows of every rank == single-GPU rollout, torch.equal"}


def run_ours(args):
    import torch.distributed as dist
    import mswe_gnn_b200  # noqa: F401
    from mswe_gnn_b200 import lib
    from mswe_gnn_b200.models.gnn import MSGNN
    from mswe_gnn_b200.parallel import PartitionedRollout, partition_graph
    from mswe_gnn_b200.training.train import RolloutRunner, rollout_test
    from mswe_gnn_b200.utils.synthetic import make_tri_mesh

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib.load()
    wl = args.workload or "cfg3"
    nx, ny = mesh_for(world, wl)
    K, W = args.steps, max(args.warmup, 3)
    model = MSGNN(**CTOR).to(dev)
    partitioned = world > 1 and args.multi != "replicas"
    host = make_tri_mesh(nx, ny, S, rollout_steps=K + W, seed=0 if partitioned else rank, with_y=False)
    N_global = host.x.shape[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    parity = None
    if partitioned:
        # ONE mesh cut over the ranks (blocks of coarsest cells); halo rows stored straight into the neighbours' arrays over
        # NVLink (parallel.PeerHalo), the step captured in a CUDA graph
        parity = partition_parity_check(model, dev, world, rank)
        if args.check_full:
            host.y = torch.empty(0, 2, 2)
            parity["full_mesh"] = partition_parity_check(model, dev, world, rank, full_graph=host, steps=2)
        part = partition_graph(host, world, rank)
        runner = PartitionedRollout(model, host, K + W, dev, transport=args.transport, part=part)
        N_nodes = sum(part.n_owned)
        N_total = N_global
    else:
        data = host.to(dev)
        runner = RolloutRunner(model, data, K + W, use_cuda_graph=True)
        N_nodes = N_global
        N_total = N_global * world

    runner.run(W)                                                          # warm-up (includes graph capture)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    runner.run(K)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    per_step_launches = runner.launches_per_step
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = N_total * K / (ms * 1e-3)

    # ---- end-to-end through the public API with HOST (pinned) inputs
    src_graph = part.graph if partitioned else host
    host_p = src_graph.clone()
    for k in host_p.keys():
        v = getattr(host_p, k)
        if torch.is_tensor(v):
            setattr(host_p, k, v.pin_memory())
    host_p.y = torch.empty(0, 2, K)                     # only its last dimension (number of steps) is read
    h2d = sum(getattr(host_p, k).numel() * getattr(host_p, k).element_size() for k in host_p.keys()
              if torch.is_tensor(getattr(host_p, k)) and k != "y")
    n_out = N_nodes
    out_host = torch.empty(K, n_out, 2, dtype=torch.float32).pin_memory()

    if partitioned:
        part_p = part
        owned_rows = torch.from_numpy(part.owned_rows).to(dev)

        import copy
        pp = copy.copy(part_p)
        pp.graph = host_p                               # pinned local graph
        r_e2e = PartitionedRollout(model, None, K, dev, transport=args.transport, part=pp)

        def e2e_call():
            # the next simulation on the same partitioned mesh: inputs from pinned host memory into the runner's buffers
            # (plan, peer arena and captured step are reused, as rollout_test's runner cache does on one GPU)
            r_e2e.rebind(host_p)
            r_e2e.run(out_host=out_host)                # owned rows of every step -> pinned host while the next step runs
            torch.cuda.synchronize()
    else:
        def e2e_call():
            ta = time.perf_counter()
            g = host_p.to(dev, non_blocking=True)
            torch.cuda.synchronize(); tb = time.perf_counter()
            rollout_test(model, g, out_host=out_host)  # every step's predictions -> pinned host while the next step runs
            torch.cuda.synchronize(); tc_ = time.perf_counter()
            if os.environ.get("BENCH_E2E_DEBUG"):
                print(f"e2e: h2d {1e3 * (tb - ta):.1f} ms, rollout_test {1e3 * (tc_ - tb):.1f} ms, d2h {1e3 * (time.perf_counter() - tc_):.1f} ms",
                      file=sys.stderr)

    e2e_call()                                          # warm (first-touch allocations of the caching allocator)
    e2e_times = []
    for _ in range(3):                                  # median of 3 complete calls (every call rebuilds plan + graph)
        barrier()
        t0 = time.perf_counter()
        e2e_call()
        barrier()
        e2e_times.append(time.perf_counter() - t0)
    e2e_s = sorted(e2e_times)[1]
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = N_total * K / e2e_s
    if partitioned:
        r_e2e.close()
    halo_info = None
    if partitioned:
        hb = torch.tensor([float(runner.halo_bytes_per_step) / max(runner.exchanges_per_step, 1), float(sum(part.n_halo))], device=dev)
        dist.all_reduce(hb, op=dist.ReduceOp.MAX)
        halo_info = {"exchanges_per_step": int(runner.exchanges_per_step), "max_bytes_per_exchange": int(hb[0].item()),
                     "max_halo_nodes_per_rank": int(hb[1].item())}

    # ---- per-kernel timing of one eager (non-graph) step with CUDA events (roofline of the dominant kernel);
    # a partitioned step exchanges halos, so every rank has to take part
    alg = algorithmic_bytes(nx, ny)
    if partitioned:                                      # per-rank share of the algorithmic bytes
        alg = {k: (v / world if not isinstance(v, list) else [x / world for x in v]) for k, v in alg.items()}
    kern = profile_kernels(runner, alg) if (partitioned or rank == 0) else None
    if rank != 0:
        if world > 1:
            if not args.no_training_extra:
                training_extra(world, rank, dev)
            if partitioned:
                runner.close()
            dist.destroy_process_group()
        return
    pk = peaks()
    dom = max(kern.values(), key=lambda r: r["ms_per_step"])
    try:                                                 # measured DRAM bytes of one level-0 launch (ncu --set full capture)
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic_r02.json"))).get(dom["name"])
        traffic = {"dram_bytes_per_level0_launch": tr["dram_bytes_per_launch"],
                   "algorithmic_bytes_per_level0_launch": tr["algorithmic_bytes_per_launch"], "capture": tr["capture"]} if tr else None
    except Exception:
        traffic = None
    if dom["bound"] == "hbm":
        roof = {"bound": "hbm", "kernel": dom["name"], "achieved": dom["gbs"], "peak": pk["hbm"], "unit": "GB/s",
                "frac": dom["gbs"] / pk["hbm"], "traffic": traffic, "peak_source": pk["hbm_src"]}
    else:
        roof = {"bound": "tensor", "kernel": dom["name"], "achieved": dom["tflops"], "peak": pk["bf16"], "unit": "TFLOP/s",
                "frac": dom["tflops"] / pk["bf16"], "traffic": traffic, "peak_source": pk["hbm_src"] + " (sustained bf16)",
                "note": ("algorithmic FLOPs of the reference edge MLP / time; executed as 3 MMAs per product (hi/lo splits for fp32 "
                         "parity): fp16 hi/lo on kind::f16 = 108 instructions of 64 cycles per 128 edges "
                         "(tools/microbench/mma_rate3.cu), i.e. a ceiling of ~0.50 of the bf16 peak for this formulation; the "
                         "3xTF32 kernel (MSWE_GATE=tc) needs 216 instructions: ~0.25")}
    step_gbs = alg["total"] / (ms * 1e-3 / K) / 1e9

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_rate(1, 1, wl=wl if wl != "cfg4" else "cfg3")
        cpu = {"value": r["value"], "unit": "node-steps/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    training = None
    if not args.no_training_extra:
        training = training_extra(world, rank, dev)
    if partitioned:
        multi = (f"ONE mesh tri({nx},{ny}) = {N_global} nodes cut over {world} GPUs by blocks of coarsest cells "
                 f"(every level sharded, {halo_info['exchanges_per_step']} halo exchanges per step, "
                 f"<= {halo_info['max_bytes_per_exchange']} B each, transport {args.transport}: " +
                 ("boundary rows stored into the neighbours' IPC-mapped arrays by one kernel per exchange, device-side flags" if args.transport == "peer" else "NCCL send/recv") +
                 "); " + ("per-GPU share fixed at 2 x cfg3's nodes (weak scaling; N=1 is cfg3 itself)" if wl == "cfg3" else "fixed total mesh (strong scaling)"))
    elif world > 1:
        multi = "independent simulations per rank (replicas, no collective)"
    else:
        multi = "single GPU"
    line = {"metric": "mSWE-GNN rollout node-steps/sec", "value": value, "unit": "node-steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "strong" if (partitioned and wl != "cfg3") else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{wl}: default config.yaml mSWE-GNN (K=4,F=64,mlp_layers=3,S=4) autoregressive rollout on "
                                   f"tri({nx},{ny}) = {N_global} nodes ({N_nodes} owned per GPU); random-init weights seed 666; 30% wet nodes",
                       "l2": f"working set {alg['total'] / 1e9:.1f} GB per step per GPU >> 126 MB L2 (inputs larger than L2, no flush needed)",
                       "multi_gpu": multi, "cuda_graph": bool(getattr(runner, "use_cuda_graph", False))},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "node-steps/s", "h2d_bytes_per_step": h2d // K, "d2h_bytes_per_step": n_out * 8,
                    "what": ("PartitionedRollout.rebind(pinned local graph) + run(): node inputs, boundary series and edge attributes "
                             "host -> device, K captured steps with halo exchange, owned predictions of every step -> pinned host on a side "
                             "stream while the next step runs (run(out_host=...)); partitioning, plan and peer arena built once per "
                             "mesh; median of 3 calls") if partitioned else
                            "rollout_test(model, graph, out_host=pinned): pinned host graph -> device, K steps, every step's predictions -> pinned "
                            "host on a side stream while the next step runs; the runner "
                            "(plan, workspaces, captured step) is cached per mesh under a content hash of the topology; median of 3 calls"},
            "gpu_launches": per_step_launches * K,
            "roofline": roof,
            "hbm_fraction_whole_step": {"algorithmic_GB_per_step": alg["total"] / 1e9, "achieved_GBps": step_gbs,
                                        "frac_of_peak": step_gbs / pk["hbm"], "peak_source": pk["hbm_src"]},
            "kernels": sorted(kern.values(), key=lambda r: -r["ms_per_step"]),
            "halo": halo_info,
            "parity_check": parity,
            "training": training,
            "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        if partitioned:
            runner.close()
        dist.destroy_process_group()



# --------------------------------------------------------------------------------------------------
# training workloads (secondary metric: BASELINE.json configs[1] and configs[4])
# --------------------------------------------------------------------------------------------------
TRAIN_WORKLOADS = {
    # name: (model, (nx, ny), graphs per rank, rollout steps)
    "cfg2-train": ("GNN", (160, 160), 8, 1),      # single-scale SWE-GNN, 8 x 51,201-node meshes, 1 GPU
    "cfg5-train": ("MSGNN", (224, 224), 4, 1),    # data parallel: 4 x 133,284-node simulations per GPU, gradient all-reduce
}


def _train_setup(wl, rank, dev, small=False):
    from mswe_gnn_b200.models.gnn import GNN, MSGNN
    from mswe_gnn_b200.utils.data import Batch
    from mswe_gnn_b200.utils.synthetic import make_single_scale_mesh, make_tri_mesh
    kind, (nx, ny), G, R = TRAIN_WORKLOADS[wl]
    if small:
        nx, ny, G = 64, 64, 1
    if kind == "GNN":
        cfg = {k: v for k, v in MODEL_CFG.items() if k not in ("learned_pooling", "skip_connections")}
        ctor = dict(num_node_features=8, num_edge_features=1, previous_t=3, n_GNN_layers=2, **cfg)
        graphs = [make_single_scale_mesh(nx, ny, rollout_steps=R, seed=1000 * rank + i) for i in range(G)]
        model = GNN(**ctor)
    else:
        ctor = CTOR
        graphs = [make_tri_mesh(nx, ny, S, rollout_steps=R, seed=1000 * rank + i) for i in range(G)]
        model = MSGNN(**ctor)
    batch = Batch.from_data_list(graphs)
    return kind, ctor, model.to(dev) if dev is not None else model, batch, R, (nx, ny, G)


def cpu_training_rate(wl, threads=None):
    """Reference training step (oracle port of training/train.py:125-145 + torch.autograd) on the host cores."""
    from oracle import swe_oracle as O
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    kind, ctor, model, batch, R, (nx, ny, G) = _train_setup(wl, 0, None, small=True)
    g = batch._graphs[0]
    spec = O.ModelSpec(kind, **ctor)
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    times = []
    for _ in range(2):
        t0 = time.perf_counter()
        loss = O.training_step(sd, spec, g, R)
        loss.backward()
        times.append(time.perf_counter() - t0)
    n = g.x.shape[0] * R
    return dict(value=n / times[-1], cores=threads,
                sample=f"1 training step (fwd+bwd, {R} rollout step) of the same {kind} on one tri({nx},{ny}) graph ({g.x.shape[0]} nodes), "
                       f"torch {torch.__version__} CPU autograd, {threads} threads")


def training_extra(world, rank, dev, steps=10, warmup=3):
    """Secondary metric on the default line (driver-visible): the training step.  N=1: cfg2-train (BASELINE.json configs[1],
    single-scale SWE-GNN, 8 x 51,201-node graphs) as one captured graph; N>1: cfg5-train (configs[4]) data parallel over
    simulations, 4 x 133,284-node graphs per GPU, one all-reduce of the flat gradient per step."""
    import torch.distributed as dist
    from mswe_gnn_b200.training.optim import FlatAdamW
    from mswe_gnn_b200.training.train import TrainStepRunner, training_step
    wl = "cfg2-train" if world == 1 else "cfg5-train"
    kind, ctor, model, batch_host, R, (nx, ny, G) = _train_setup(wl, rank, dev)
    batch = batch_host.to(dev)
    opt = FlatAdamW(model, lr=3e-3, weight_decay=0.0, max_norm=1.0)
    runner = TrainStepRunner(model, batch, opt, rollout_steps=R, use_cuda_graph=(world == 1), warmup=warmup)
    step = (lambda: runner.step()) if world == 1 else \
        (lambda: training_step(model, batch, R, only_where_water=True, velocity_scaler=7.0, optimizer=opt))
    from mswe_gnn_b200 import lib as _lib
    c0 = _lib.launch_count
    for _ in range(warmup if world > 1 else 1):
        step()
    launches = (_lib.launch_count - c0) // (warmup if world > 1 else 1) if world > 1 else runner.launches_per_step
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    n = int(batch.x.shape[0])
    return {"metric": "mSWE-GNN training node-steps/sec (forward+backward+clip+AdamW)", "workload": wl, "value": n * R * world * steps / (ms * 1e-3),
            "unit": "node-steps/s", "ms_per_step": ms / steps, "n_gpus": world, "steps": steps, "nodes_per_gpu": n,
            "cuda_graph": runner._graph is not None, "launches_per_step": launches,
            "parallelism": "single GPU" if world == 1 else f"data parallel over simulations (dp{world}), all-reduce of the flat fp32 gradient"}


def run_train(args):
    import torch.distributed as dist
    import mswe_gnn_b200  # noqa: F401
    from mswe_gnn_b200 import lib
    from mswe_gnn_b200.training.train import training_step
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib.load()
    wl = args.workload
    kind, ctor, model, batch_host, R, (nx, ny, G) = _train_setup(wl, rank, dev)
    for k in batch_host.keys():
        v = getattr(batch_host, k)
        if torch.is_tensor(v):
            setattr(batch_host, k, v.pin_memory())
    from mswe_gnn_b200.training.optim import FlatAdamW
    from mswe_gnn_b200.training.train import TrainStepRunner
    batch = batch_host.to(dev)
    # config.yaml lr_info; gradient_clip_val = 1 (main.py:109); loss, clip and AdamW are device kernels on flat buffers
    opt = FlatAdamW(model, lr=3e-3, weight_decay=0.0, max_norm=1.0)
    n_nodes = batch.x.shape[0]
    K, W = args.steps, max(args.warmup, 3)
    # one GPU: the whole step (forward, loss, backward, clip, AdamW) is ONE captured CUDA graph; data parallel: eager, with
    # the all-reduce of the flat gradient between backward and update
    runner = TrainStepRunner(model, batch, opt, rollout_steps=R, use_cuda_graph=(world == 1 and not args.no_graph))

    def step(b):
        if world == 1:
            return runner.step(b.x, b.y, b.BC) if b is not batch else runner.step()
        return training_step(model, b, R, only_where_water=True, velocity_scaler=7.0, optimizer=opt)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step(batch)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(K):
        step(batch)
    ev1.record()
    barrier()
    launches = runner.launches_per_step * K
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    # e2e: the batch comes from pinned host memory every step, the loss goes back to the host
    h2d = sum(getattr(batch_host, k).numel() * getattr(batch_host, k).element_size() for k in batch_host.keys()
              if torch.is_tensor(getattr(batch_host, k)))
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        loss_host = float(step(batch_host.to(dev, non_blocking=True)) if world > 1 else
                          runner.step(batch_host.x, batch_host.y, batch_host.BC))
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([ms, e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1])
    value = n_nodes * R * world * K / (ms * 1e-3)
    kern = profile_train_kernels(lambda: training_step(model, batch, R, only_where_water=True, velocity_scaler=7.0, optimizer=opt))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    dom = max(kern.values(), key=lambda r: r["ms_per_step"])
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_training_rate(wl)
        cpu = {"value": r["value"], "unit": "node-steps/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}
    line = {"metric": "mSWE-GNN training node-steps/sec (forward+backward+AdamW)", "value": value, "unit": "node-steps/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{wl}: {kind} (config.yaml hyper-parameters: F=64, K=4, mlp_layers=3) training step, {G} x tri({nx},{ny}) "
                                   f"graphs per GPU = {n_nodes} nodes, {R} rollout step(s), loss RMSE on wet cells, grad-clip 1, AdamW "
                                   f"(device kernels on flat buffers); " + ("whole step = one captured CUDA graph" if runner._graph is not None else "eager step"),
                       "l2": "saved activations >> 126 MB L2 (inputs larger than L2, no flush needed)",
                       "multi_gpu": "data parallel over simulations, one all-reduce of the flat fp32 gradient per step" if world > 1 else "single GPU"},
            "clocks": clocks,
            "e2e": {"value": n_nodes * R * world * K / e2e_s, "unit": "node-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "what": "step(x, y, BC of the batch from pinned host memory -> captured buffers), loss.item() back, every step"},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": dom["name"], "achieved": dom["tflops"], "peak": pk["bf16"], "unit": "TFLOP/s",
                         "frac": dom["tflops"] / pk["bf16"], "traffic": _traffic_of(dom["name"]),
                         "peak_source": pk["hbm_src"] + " (sustained bf16)",
                         "note": "algorithmic FLOPs / time of the entry point with the largest share of the step; the wide edge-MLP "
                                 "GEMMs run as 3xTF32 on tcgen05 (fp32-accurate), the narrow ones in exact fp32 on CUDA cores"},
            "kernels": sorted(kern.values(), key=lambda r: -r["ms_per_step"])[:12],
            "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _traffic_of(name):
    """Measured DRAM bytes of one launch of `name` from the committed ncu --set full captures (or None)."""
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic_r02.json"))).get(name)
        return {"dram_bytes_per_launch": tr["dram_bytes_per_launch"],
                "algorithmic_bytes_per_launch": tr["algorithmic_bytes_per_launch"], "capture": tr["capture"]} if tr else None
    except Exception:
        return None


def profile_train_kernels(step_fn):
    """CUDA-event pair around every C-ABI launch of one training step, grouped by entry point."""
    from mswe_gnn_b200 import lib
    records, orig = [], {}
    names = [n[4:] for n in lib.SIGNATURES if n.startswith("swe_") and hasattr(lib, n[4:]) and callable(getattr(lib, n[4:]))
             and n[4:] not in ("mlp_layer_bwd_dx_grid", "mlp_layer_bwd_dw_grid", "mlp_layer_bwd_dw_tc_grid", "mlp_layer_bwd_dx_tc_grid", "gate_tc_image_bytes", "hop_tc_image_bytes", "csr_build")]

    def wrap(name):
        fn = getattr(lib, name)
        orig[name] = fn

        def inner(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            fl = 0.0
            if name == "mlp_layer_bwd_dx" and a[11] is not None:
                fl = 2.0 * a[4] * a[5] * a[10]
            elif name == "mlp_layer_bwd_dw":
                fl = 2.0 * a[1] * a[2] * a[4]
            elif name == "mlp_layer_fwd":
                fl = 2.0 * a[1] * a[4] * sum(a[0].seg[j].width for j in range(a[0].n_seg))
            elif name == "mlp_layer_bwd_dx_tc_fused":
                fl = 2.0 * a[4] * a[5] * a[10]
            elif name == "mlp_layer_bwd_dx_tc":
                fl = 2.0 * a[1] * a[2] * a[7]
            elif name == "mlp_layer_bwd_dw_tc":
                fl = 2.0 * a[1] * a[2] * sum(a[3].seg[j].width for j in range(a[3].n_seg))
            elif name == "edge_gate_tc_train_fwd":
                fl = 2.0 * a[6] * (a[8] * 128 + 128 * 128 + 128 * 64)
            records.append((name, e0, e1, fl))
            return r
        setattr(lib, name, inner)

    for n in names:
        wrap(n)
    try:
        step_fn()
        torch.cuda.synchronize()
    finally:
        for n, fn in orig.items():
            setattr(lib, n, fn)
    out = {}
    for name, e0, e1, fl in records:
        r = out.setdefault(name, dict(name="swe_" + name, ms_per_step=0.0, launches_per_step=0, flops_per_step=0.0))
        r["ms_per_step"] += e0.elapsed_time(e1)
        r["launches_per_step"] += 1
        r["flops_per_step"] += fl
    tot = sum(r["ms_per_step"] for r in out.values())
    for r in out.values():
        r["tflops"] = r["flops_per_step"] / (r["ms_per_step"] * 1e-3) / 1e12 if r["ms_per_step"] > 0 else 0.0
        r["share"] = r["ms_per_step"] / tot if tot else 0.0
    return out


def profile_kernels(runner, alg):
    """One eager (non-graph) step with a CUDA-event pair around every C-ABI launch, grouped by
    kernel.  Events are recorded on torch's current stream, the stream the kernels launch on."""
    from mswe_gnn_b200 import lib
    records = []
    orig = {}
    names = ["row_mlp_tc", "row_mlp_tc16", "row_linear_tc16", "node_encode_fwd", "edge_gate_fwd", "edge_gate_tc_fwd", "edge_gate_tc16_fwd", "propagate_hop_tc16_fwd", "propagate_hop_tc16s_fwd", "edge_gate_tc_dec_fwd", "gate_partials_tc", "edge_gate_tc_stat_fwd",
             "gate_static_partials_tc", "node_linear_fwd", "propagate_hop_fwd", "propagate_hop_tc_fwd", "pool_mean_fwd",
             "decode_head_fwd", "edge_encode_fwd", "apply_bc", "step_advance", "halo_exchange", "pack_rows"]

    def wrap(name):
        fn = getattr(lib, name)
        orig[name] = fn

        def inner(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            meta = None
            if name in ("propagate_hop_fwd", "propagate_hop_tc_fwd", "propagate_hop_tc16_fwd", "propagate_hop_tc16s_fwd"):
                meta = ("hop", int(a[6]), a[7] is not None)          # n_dst, has filter
            elif name in ("edge_gate_fwd", "edge_gate_tc_fwd", "edge_gate_tc16_fwd"):
                meta = ("gate", int(a[6]), a[3] is not None)         # n_edges, has edge features
            elif name == "edge_gate_tc_dec_fwd":
                meta = ("gate", int(a[5]), a[2] is not None)
            elif name == "edge_gate_tc_stat_fwd":                    # reference work of the call: the whole edge MLP
                meta = ("gate", int(a[6]), int(a[8]) == 5 * F)
            records.append((name, e0, e1, meta))
            return r
        setattr(lib, name, inner)

    for n in names:
        wrap(n)
    try:
        reps = 3
        runner.reset()                                   # step counter back to 0: stay inside the BC / prediction buffers
        for _ in range(reps):
            runner._one_step()
        torch.cuda.synchronize()
    finally:
        for n, fn in orig.items():
            setattr(lib, n, fn)
    out = {}
    N, E = alg["N"], alg["E"]
    for name, e0, e1, meta in records:
        r = out.setdefault(name, dict(name="swe_" + name, ms_per_step=0.0, launches_per_step=0, bytes_per_step=0.0,
                                      flops_per_step=0.0, bound="hbm"))
        r["ms_per_step"] += e0.elapsed_time(e1) / reps
        r["launches_per_step"] += 1.0 / reps
        if meta and meta[0] == "hop" and meta[2]:
            s = min(range(S), key=lambda i: abs(N[i] - meta[1]))
            r["bytes_per_step"] += alg["hop"][s] / reps
        if meta and meta[0] == "gate":
            nseg = (5 if meta[2] else 3)
            r["flops_per_step"] += 2.0 * meta[1] * (nseg * F * 2 * F + 2 * F * 2 * F + 2 * F * F) / reps
            r["bound"] = "tensor"
    for r in out.values():
        sec = r["ms_per_step"] * 1e-3
        r["gbs"] = r["bytes_per_step"] / sec / 1e9 if sec > 0 else 0.0
        r["tflops"] = r["flops_per_step"] / sec / 1e12 if sec > 0 else 0.0
    tot = sum(r["ms_per_step"] for r in out.values())
    for r in out.values():
        r["share"] = r["ms_per_step"] / tot if tot else 0.0
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=48,
                    help="timed rollout steps (default 48 = the reference's rollout length: 96 h at 120 min, train.py:84)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=[None, *WORKLOADS, *TRAIN_WORKLOADS])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="training workloads: eager step instead of the captured graph")
    ap.add_argument("--no-training-extra", action="store_true", help="default workload: skip the secondary training measurement")
    ap.add_argument("--transport", default="peer", choices=["peer", "nccl"],
                    help="N>1 halo exchange: peer-memory stores + device flags in a captured step (default) or NCCL send/recv (eager)")
    ap.add_argument("--check-full", action="store_true",
                    help="N>1: also compare the partitioned rollout of the benchmark mesh itself with rank 0's single-GPU rollout")
    ap.add_argument("--multi", default="partitioned", choices=["partitioned", "replicas"],
                    help="N>1: one mesh partitioned over the GPUs with halo exchange (default) or independent replicas")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload in TRAIN_WORKLOADS:
        run_train(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
