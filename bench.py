"""Benchmark of the FuMI episodic inner-loop path (BASELINE.json: meta-test episodes/s and meta-train tasks/s).

    python bench.py --gpus N --steps K --warmup W           our arm (one process per GPU; torchrun for N>1)
    python bench.py --impl reference --steps K --warmup W   the reference's CPU path on the host cores

The JSON line's headline (`metric`, `value`, `e2e`, `roofline`) is BASELINE.json configs[1]: FuMI 5-way 5-shot
meta-train, 5 inner steps, 4096 tasks per meta-batch per GPU (query 32/class, dropout 0.25, Adam lr 3e-5 wd 5e-4:
the reference defaults, utils.py:19-229), on the synthetic iNat-Anim-shaped bank.  A "step" = one
meta-batch: hypernetwork, first-layer projection of the split's bank, Gram blocks of every task, fused inner loop +
query scoring, second-order backward, dW0, (all-reduce), Adam.  The same line carries every other BASELINE config
under `secondary` (5w5s / 5w1s meta-test, MAML meta-train, AM3 10-way meta-test, 20-way 10-step meta-train), each
with value, ms_per_step, e2e, the whole-step roofline fraction and a CPU baseline; `--workload X` runs X alone.

value = whole-job tasks/s with the sampled indices already resident in HBM; e2e = the same through the public API
(native host sampler -> pinned plan -> H2D -> evaluate -> D2H of loss/acc) inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "fumi_train_5w5s": dict(model="fumi", N=5, K=5, Q=32, steps=5, train=True, tasks=4096),      # BASELINE configs[1]
    "maml_train_5w5s": dict(model="maml", N=5, K=5, Q=32, steps=5, train=True, tasks=4096),      # configs[2]
    "fumi_test_5w5s": dict(model="fumi", N=5, K=5, Q=20, steps=100, train=False, tasks=4096),    # headline meta-test
    "fumi_test_5w1s": dict(model="fumi", N=5, K=1, Q=20, steps=100, train=False, tasks=4096),    # configs[0]
    "am3_test_10w5s": dict(model="am3", N=10, K=5, Q=10, steps=0, train=False, tasks=4096),      # configs[3]
    "fumi_train_20w5s": dict(model="fumi", N=20, K=5, Q=32, steps=10, train=True, tasks=1024),   # configs[4]
}
PRIMARY = "fumi_train_5w5s"
SECONDARY = ["fumi_test_5w5s", "fumi_test_5w1s", "maml_train_5w5s", "am3_test_10w5s", "fumi_train_20w5s"]


def metric_of(wl):
    w = WORKLOADS[wl]
    name = {"fumi": "FuMI", "maml": "MAML", "am3": "AM3"}[w["model"]]
    if w["model"] == "am3":
        return f"meta-test episodes/sec ({name} {w['N']}-way {w['K']}-shot, prototype + text mixing)", "episodes/s"
    kind = "meta-train tasks/sec" if w["train"] else "meta-test episodes/sec"
    return f"{kind} ({name} {w['N']}-way {w['K']}-shot, {w['steps']} inner steps)", ("tasks/s" if w["train"] else "episodes/s")


def algorithmic_bytes(N, K, Q, D, T, train):
    """B_fwd = 4(NK+NQ)D + 4NT + 8(NK+NQ) + 4 NQ N + 8 NQ ; B_train = B_fwd + 4(NK+NQ)D."""
    NK, NQ = N * K, N * Q
    b = 4 * (NK + NQ) * D + 4 * N * T + 8 * (NK + NQ) + 4 * NQ * N + 8 * NQ
    return b + (4 * (NK + NQ) * D if train else 0)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons of one GPU sampled DURING the timed region (B200_PROFILING.md clocks line):
    NVML polled every 5 ms from a thread when pynvml is importable, else `nvidia-smi -lms 100`."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.nvml, self.samples, self.stop_flag = None, [], threading.Event()

    def _poll(self):
        n = self.nvml
        try:
            h = n.nvmlDeviceGetHandleByIndex(self._nvml_index())
            mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
            while not self.stop_flag.is_set():
                self.samples.append((n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM), mx,
                                     n.nvmlDeviceGetCurrentClocksEventReasons(h)))
                time.sleep(0.005)
        except Exception as e:          # fall through to whatever was collected
            self.err = repr(e)

    def _nvml_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.nvml is not None:
            self.stop_flag.set()
            self.thread.join(timeout=1.0)
            n = self.nvml
            bits = {"hw_slowdown": n.nvmlClocksEventReasonHwSlowdown,
                    "hw_thermal_slowdown": n.nvmlClocksEventReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": n.nvmlClocksEventReasonSwThermalSlowdown,
                    "sw_power_cap": n.nvmlClocksEventReasonSwPowerCap}
            sm = [x[0] for x in self.samples]
            reasons = sorted(nm for nm, bit in bits.items() if any(x[2] & bit for x in self.samples))
            return {"sm_mhz": float(np.median(sm)) if sm else None,
                    "sm_max_mhz": float(self.samples[0][1]) if sm else None, "samples": len(sm), "reasons": reasons,
                    "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}



def make_args(model, N, K, Q, steps, train, device, D, T, tasks, dropout):
    from fumi_b200 import utils
    a = utils.parser().parse_args(["--model", model, "--num_ways", str(N), "--num_shots", str(K),
                                   "--num_shots_test", str(Q), "--batch_size", str(tasks),
                                   "--num_train_adapt_steps", str(steps), "--num_test_adapt_steps", str(steps),
                                   "--im_emb_dim", str(D), "--text_emb_dim", str(T), "--dropout", str(dropout),
                                   "--synthetic", "--wandb_offline"])
    a.device = device
    return a


def cpu_model_name():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


# ------------------------------------------------------------------------------------------ CPU leg
_HOST_BANK = {}


def host_bank(images, classes, D, T):
    from fumi_b200.data.synth import make_bank
    key = (images, classes, D, T)
    if key not in _HOST_BANK:
        _HOST_BANK[key] = make_bank(num_images=images, num_classes=classes, im_dim=D, text_dim=T)
    return _HOST_BANK[key]


def cpu_reference_leg(wl, a, min_steps, warmup, budget_s, min_s=0.0, tasks_per_batch=4, seed=123):
    """The reference's CPU path for the workload's episode shape, on the benched bank, with the reference defaults
    (--batch_size 4, --dropout 0.25 in FuMI's train mode, Adam lr 3e-5 wd 5e-4), all host threads.  The reference is
    pure Python and cannot travel to the GPU box, so this is its port (kind='port'):
    oracle/sampler_np.FlatSampler (the loader: random / numpy / torch generators), oracle/episode_torch.py (autograd
    restatement of fumi.py:148-193 / maml.py:158-191, per-task Python loop) or oracle/episode_np.am3_batch
    (am3.py:128-212), plain torch.nn.Linear layers in the reference's construction order, torch.optim.Adam.  Nothing of
    libfumi_b200.so is loaded by this leg."""
    import random
    import torch
    import torch.nn as nn
    from oracle import episode_np, episode_torch, sampler_np
    from fumi_b200.data.synth import class_split
    w = WORKLOADS[wl]
    model_name, N, K, Q, steps, train = w["model"], w["N"], w["K"], w["Q"], w["steps"], w["train"]
    D, T = a.im_dim, a.text_dim
    torch.set_num_threads(os.cpu_count())
    bank = host_bank(a.bank_images, a.bank_classes, D, T)
    cats = class_split(a.bank_classes)[0 if train else 2]
    sampler = sampler_np.FlatSampler(sampler_np.class_tables(bank.cat_of, cats), N, K, Q)
    torch.manual_seed(seed); np.random.seed(seed); random.seed(seed)
    if model_name == "fumi":        # fumi.py:70-107: hyper_net Linear(T,256)-ReLU-Linear(256,65); im_net linear0, linear1
        layers = {"hyper_net.0": nn.Linear(T, 256), "hyper_net.2": nn.Linear(256, 65),
                  "im_net.linear0": nn.Linear(D, 256), "im_net.linear1": nn.Linear(256, 64)}
    elif model_name == "maml":      # maml.py:21-29
        layers = {"net.lin_0": nn.Linear(D, 256), "net.lin_1": nn.Linear(256, 64), "net.lin_final": nn.Linear(64, N)}
    else:                           # am3.py:16-88 (prototype_dim 64, text_hid_dim 256)
        layers = {"image_encoder": nn.Linear(D, 64), "g.0": nn.Linear(T, 256), "g.3": nn.Linear(256, 64),
                  "h.0": nn.Linear(64, 256), "h.3": nn.Linear(256, 1)}
    params = {f"{k}.{n}": p for k, l in layers.items() for n, p in l.named_parameters()}
    opt = torch.optim.Adam(list(params.values()), lr=3e-5, weight_decay=5e-4) if train else None
    feats, text = torch.from_numpy(bank.feats), torch.from_numpy(bank.text)
    dropout = float(a.dropout) if (train and model_name == "fumi") else 0.0
    sampler.new_iterator()

    def one_batch():
        b = sampler.next_batch(tasks_per_batch)
        t = lambda x: torch.from_numpy(x)
        sup_ids, qry_ids = t(b["sup_ids"]), t(b["qry_ids"])
        cls_of_label = np.empty_like(b["classes"])                      # split-class carrying label i (fumi.py:207-210)
        np.put_along_axis(cls_of_label, b["label_perm"], b["classes"], axis=1)
        batch = dict(sup_x=feats[sup_ids], qry_x=feats[qry_ids], sup_y=t(b["sup_targets"]), qry_y=t(b["qry_targets"]),
                     class_text=text[t(cats[cls_of_label])])
        if model_name == "fumi":
            episode_torch.fumi_batch(params, batch, 0.01, steps, train=train, dropout_p=dropout)
        elif model_name == "maml":
            episode_torch.maml_batch(params, batch, 0.01, steps, train=train)
        else:
            with torch.no_grad():
                npb = dict(sup_x=batch["sup_x"].numpy(), qry_x=batch["qry_x"].numpy(), sup_y=b["sup_targets"],
                           qry_y=b["qry_targets"], sup_text=bank.text[bank.cat_of[b["sup_ids"]]])
                episode_np.am3_batch({k: v.detach().numpy() for k, v in params.items()}, npb, N)
        if train:
            opt.step()

    for _ in range(warmup):
        one_batch()
    per_step, t_all0 = [], time.perf_counter()
    while True:
        t0 = time.perf_counter()
        one_batch()
        per_step.append(time.perf_counter() - t0)
        el = time.perf_counter() - t_all0
        if el > budget_s or (len(per_step) >= min_steps and el >= min_s):
            break
    tot = sum(per_step)
    unit = metric_of(wl)[1]
    return dict(value=tasks_per_batch * len(per_step) / tot, unit=unit, cores=torch.get_num_threads(), kind="port",
                cpu=cpu_model_name(),
                sample=f"{len(per_step)} meta-batches x {tasks_per_batch} tasks (reference --batch_size 4) in {tot:.1f} s, {wl}, "
                       f"D={D}, bank {a.bank_images} x {D}, dropout {dropout}, loader oracle/sampler_np + "
                       f"{'oracle/episode_np.am3_batch' if model_name == 'am3' else 'autograd port oracle/episode_torch.py'}"
                       f"{', torch Adam' if train else ''}",
                ms_per_step=1e3 * tot / len(per_step), steps_done=len(per_step))


# ------------------------------------------------------------------------------------------ our arm
def ncu_traffic(kernel, workload, tasks):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the newest committed `ncu --set full`
    summary taken on this workload at this batch size (else None)."""
    pdir = os.path.join(ROOT, "profiles")
    for name in sorted((f for f in os.listdir(pdir) if f.endswith("_summary.json") and "ncu_full" in f), reverse=True):
        try:
            with open(os.path.join(pdir, name)) as f:
                d = json.load(f)
            if d["workload"] != workload or int(d["tasks"]) != int(tasks):
                continue
            recs = d["kernels"][kernel]
            return max(r["traffic_bytes"] for r in recs), "profiles/" + name
        except (OSError, KeyError, ValueError):
            continue
    return None, None


class Context:
    """Per-process state shared by the workloads: device, the host bank, one FeatureBank per split."""

    def __init__(self, a):
        import torch
        import torch.distributed as dist
        self.a = a
        self.rank = int(os.environ.get("RANK", 0))
        self.world = int(os.environ.get("WORLD_SIZE", 1))
        self.local = int(os.environ.get("LOCAL_RANK", 0))
        torch.cuda.set_device(self.local)
        self.device = torch.device("cuda", self.local)
        if self.world > 1 and not dist.is_initialized():
            dist.init_process_group("nccl", device_id=self.device)
        t0 = time.time()
        self.bank = host_bank(a.bank_images, a.bank_classes, a.im_dim, a.text_dim)
        self.setup_s = time.time() - t0
        self._fb = {}

    def split_bank(self, split, sampler):
        """FeatureBank of a split (rows in the order of the sampler's class table: identical for every (N, K, Q))."""
        import torch
        from fumi_b200.data.bank import FeatureBank
        from fumi_b200.data.synth import class_split
        if split not in self._fb:
            cats = class_split(self.a.bank_classes)[split]
            self._fb[split] = FeatureBank(feats=torch.from_numpy(self.bank.feats[sampler.ids]).to(self.device),
                                          text=torch.from_numpy(self.bank.text[cats]).to(self.device),
                                          ids=sampler.ids, categories=cats)
        return self._fb[split]

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        import torch
        import torch.distributed as dist
        t = torch.tensor([ms], device=self.device)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())


def run_workload(ctx, wl, n_steps, n_warmup, kernel_pass, cpu_budget_s, parity, tasks_override=0):
    """One workload on this rank's GPU: device-resident leg (value), end-to-end leg through the public API (e2e),
    per-kernel pass, CPU baseline (rank 0, N=1)."""
    import random
    import torch
    import torch.distributed as dist
    from fumi_b200 import maml as maml_mod, utils
    from fumi_b200.data.loader import EpisodeLoader
    from fumi_b200.data.synth import class_split
    from fumi_b200.sampler import EpisodeSampler
    a, device, rank, world = ctx.a, ctx.device, ctx.rank, ctx.world
    w = WORKLOADS[wl]
    model_name, N, K, Q, steps, train = w["model"], w["N"], w["K"], w["Q"], w["steps"], w["train"]
    tasks = tasks_override or (a.tasks if (a.tasks and wl == a.workload) else w["tasks"])
    split = 0 if train else 2
    cats = class_split(a.bank_classes)[split]
    sampler = EpisodeSampler(ctx.bank.cat_of, cats, N, K, Q, num_threads=max(2, (os.cpu_count() or 8) // max(1, world)))
    fb = ctx.split_bank(split, sampler)
    dev_sampler = not a.host_sampler
    loader = EpisodeLoader(fb, sampler, tasks, device_sampler=dev_sampler)
    dropout = a.dropout if model_name != "maml" else 0.0
    args = make_args(model_name, N, K, Q, steps, train, device, a.im_dim, a.text_dim, tasks, dropout)
    args.first_order = False
    seed = 123 + rank                                   # ranks draw independent task streams (weak scaling)
    torch.manual_seed(123); np.random.seed(123); random.seed(123)
    model = utils.init_model(args, {})                   # same initial weights on every rank
    opt = utils.init_optim(args, model) if train else None
    torch.manual_seed(seed); random.seed(seed)
    eng = model._get_engine(device)
    eng.precision = a.precision
    sampler.new_iterator()
    task_name = "train" if train else "test"

    def run_resident(batch):
        if model_name == "fumi":
            eng.fumi_batch(model, batch, steps=steps, step_size=args.step_size, train=train)
        elif model_name == "maml":
            eng.maml_batch(model, batch, steps=steps, step_size=args.step_size, train=train, first_order=False)
        else:
            eng.am3_batch(model, batch, N)
        if train:
            opt.step()

    def run_api(batch):
        if model_name == "fumi":
            return model.evaluate(args, batch, opt, task=task_name)
        if model_name == "maml":
            return maml_mod.evaluate(args, model, batch, opt, task=task_name)
        return model.evaluate(batch=batch, optimizer=None, scheduler=None, num_ways=N, device=device, task="test")

    # ---- leg 1: device-resident inputs (value)
    batches = [loader.next_batch().to(device) for _ in range(n_warmup + n_steps)]
    torch.cuda.synchronize()
    model.train(train)
    for i in range(n_warmup):
        run_resident(batches[i])
    ctx.barrier()
    clocks = ClockSampler(ctx.local)
    clocks.start()
    l0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n_steps):
        run_resident(batches[n_warmup + i])
    e1.record()
    ctx.barrier()
    ms_value = ctx.max_over_ranks(e0.elapsed_time(e1))
    launches = eng.launches - l0 + (n_steps if train else 0)          # + one fused Adam launch per step
    clk = clocks.stop()
    ranks_in_sync = None
    if world > 1 and train:       # identical parameters on every rank after the all-reduced steps
        flat = torch.cat([p.detach().reshape(-1).double() for p in model.parameters()])
        mine = torch.stack([flat.sum(), (flat * flat).sum(), flat[::997].abs().sum()])
        allv = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        ranks_in_sync = bool(all(torch.equal(v, allv[0]) for v in allv))

    # ---- leg 2: end to end through the public API, host buffers (e2e).  Every step: the sampler's sequential
    # generator streams advance on the host (prefetch thread, 2 batches ahead) into a pinned plan, the plan
    # travels host -> device, fumi_sampler_expand builds the index arrays in HBM, evaluate() runs the step and
    # loss/acc come back.  (--host_sampler: the all-host sampler + pinned index arrays instead.)
    t_s = time.perf_counter()
    if dev_sampler:
        plan = sampler.plan(tasks, pin_memory=True)
        sampler_ms = (time.perf_counter() - t_s) * 1e3
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler.expand(plan, device)
        ev0.record()
        sampler.expand(plan, device)
        ev1.record()
        torch.cuda.synchronize()
        expand_ms = ev0.elapsed_time(ev1)
    else:
        loader.next_batch()
        sampler_ms, expand_ms = (time.perf_counter() - t_s) * 1e3, None
    e2e_loader = EpisodeLoader(fb, sampler, tasks, prefetch=2, device_sampler=dev_sampler)
    e2e_it = iter(e2e_loader)
    for i in range(max(2, n_warmup)):
        run_api(next(e2e_it))
    ctx.barrier()
    h2d = d2h = 0
    e0.record()
    for i in range(n_steps):
        b = next(e2e_it)
        if i == 0:
            if dev_sampler:      # the plan: classes, label_perm, head_class (i64), perm_seed (u32), picks, job_order (i32)
                h2d = tasks * N * (3 * 8 + 4 + 4 * (K + Q) + 4)
            else:
                h2d = sum(t.numel() * 8 for t in (b.sup_rows, b.qry_rows, b.sup_y, b.qry_y, b.head_class))
            # loss + acc (fumi.py:195-196); AM3's evaluate returns predictions / ids / lamdas on the host (am3.py:203-210)
            d2h = 8 if model_name != "am3" else tasks * (N * Q * 8 * 3 + N * K * 12) + 8
        run_api(b)
    e1.record()
    ctx.barrier()
    e2e_loader.close()
    ms_e2e = ctx.max_over_ranks(e0.elapsed_time(e1))

    # ---- per-kernel pass (CUDA events around every C-ABI call, on the launching stream)
    kernels, roof = {}, None
    peak, peak_src = peaks()
    bytes_task = algorithmic_bytes(N, K, Q, a.im_dim, a.text_dim, train)
    step_gbs = bytes_task * tasks / (ms_value / n_steps * 1e-3) / 1e9
    roof = {"bound": "hbm", "achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak,
            "peak_source": peak_src, "algorithmic_bytes_per_task": bytes_task,
            "note": "whole step: SURVEY 8(d) algorithmic bytes/task x tasks / measured ms_per_step (every kernel of the "
                    "meta-batch, not only the dominant one)"}
    if kernel_pass:
        eng.profile = {}
        nprof = min(3, n_steps)
        for i in range(nprof):
            run_resident(batches[n_warmup + i])
        torch.cuda.synchronize()
        for name, evs in eng.profile.items():
            ts = [x.elapsed_time(y) for x, y in evs]
            cps = max(1, len(ts) // nprof)                  # calls per step; per-step totals, median over the steps
            per_step = sorted(sum(ts[i * cps:(i + 1) * cps]) for i in range(len(ts) // cps))
            kernels[name] = {"ms_per_step": per_step[len(per_step) // 2], "calls_per_step": len(ts) / nprof}
        eng.profile = None
        top = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
        dur = kernels[top]["ms_per_step"] / kernels[top]["calls_per_step"] * 1e-3
        traffic, traffic_src = ncu_traffic(top, wl, tasks)
        roof.update({"kernel": top, "launch_ms": dur * 1e3, "traffic": traffic, "traffic_source": traffic_src,
                     "dominant_kernel": {"achieved": bytes_task * tasks / dur / 1e9,
                                         "frac": bytes_task * tasks / dur / 1e9 / peak,
                                         "note": "algorithmic bytes / launch time of the longest kernel alone"}})
        roof["by_kernel"] = {k: {"ms": v["ms_per_step"], "share_of_step": v["ms_per_step"] / (ms_value / n_steps)}
                             for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["ms_per_step"])}

    par = None
    if parity and world == 1 and model_name == "fumi" and train:
        par = parity_probe(ctx, eng, model, args, batches[n_warmup], N, K, Q, steps, dropout)

    cpu = None
    if rank == 0 and world == 1 and cpu_budget_s > 0:
        r = cpu_reference_leg(wl, a, 10 ** 9, 1, budget_s=cpu_budget_s)
        cpu = {"value": r["value"], "unit": r["unit"], "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
               "cpu": r["cpu"]}

    metric, unit = metric_of(wl)
    total = tasks * world * n_steps
    out = {"metric": metric, "value": total / (ms_value * 1e-3), "unit": unit, "ms_per_step": ms_value / n_steps,
           "steps": n_steps, "warmup": n_warmup,
           "config": {"workload": wl, "num_ways": N, "num_shots": K, "query_per_class": Q, "inner_steps": steps,
                      "tasks_per_batch_per_gpu": tasks, "im_dim": a.im_dim, "text_dim": a.text_dim, "hidden": [256, 64],
                      "parallelism": f"tasks sharded over {world} GPU(s), 1 NCCL all-reduce/step" if train else
                                     f"episodes sharded over {world} GPU(s), no collective",
                      "dropout": dropout if train else 0.0, "dense_precision": a.precision,
                      "bank_rows": int(fb.feats.shape[0]), "bank_classes": int(len(cats)),
                      "l2_policy": "per-step gathered input (tasks x rows x 8 KB) far exceeds the 126 MB L2",
                      "setup_s": round(ctx.setup_s, 1)},
           "clocks": clk,
           "e2e": {"value": total / (ms_e2e * 1e-3), "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "ms_per_step": ms_e2e / n_steps, "host_sampler_ms_per_batch": round(sampler_ms, 2),
                   "host_cores": os.cpu_count(), "device_sampler_ms_per_batch": expand_ms,
                   "path": ("EpisodeLoader(prefetch=2): fumi_sampler_plan thread -> pinned plan -> H2D -> "
                            "fumi_sampler_expand (index arrays built in HBM) -> evaluate() -> loss/acc D2H, "
                            "every step") if dev_sampler else
                           ("EpisodeLoader(prefetch=2): native sampler thread -> pinned index buffers -> H2D -> "
                            "evaluate() -> loss/acc D2H, every step")},
           "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "kernels": kernels}
    if ranks_in_sync is not None:
        out["ranks_in_sync"] = ranks_in_sync
    if par is not None:
        out["parity"] = par
    del model, opt, batches, loader, e2e_loader
    torch.cuda.empty_cache()
    return out


def parity_probe(ctx, eng, model, args, batch, N, K, Q, steps, dropout, ntasks=8):
    """The first `ntasks` tasks of a timed meta-batch, at the benched dimensions / precision / dropout, against
    oracle/episode_np in fp64 (the checker, fed the same counter-based dropout masks): query predictions must be
    equal, logits within 1e-4 and every meta-gradient within 2e-4 (tensor-normalised max error)."""
    import torch
    from oracle import episode_np
    from fumi_b200.data.bank import EpisodeBatch
    from fumi_b200.dropout import mask_array
    cut = lambda t: None if t is None else t[:ntasks].contiguous()
    sub = EpisodeBatch(bank=batch.bank, sup_rows=cut(batch.sup_rows), qry_rows=cut(batch.qry_rows), sup_y=cut(batch.sup_y),
                       qry_y=cut(batch.qry_y), sup_ids=None, qry_ids=None, head_class=cut(batch.head_class))
    model.train()
    for p in model.parameters():
        if p.grad is not None:
            p.grad.zero_()
    res = eng.fumi_batch(model, sub, steps=steps, step_size=args.step_size, train=True)
    seed = (int(getattr(model, "dropout_base_seed", 0)) << 20) + model.dropout_seed
    torch.cuda.synchronize()
    feats, text = batch.bank.feats, batch.bank.text
    NK, NQ = N * K, N * Q
    fb = dict(sup_x=feats[sub.sup_rows].double().cpu().numpy(), qry_x=feats[sub.qry_rows].double().cpu().numpy(),
              sup_y=sub.sup_y.cpu().numpy(), qry_y=sub.qry_y.cpu().numpy(),
              class_text=text[sub.head_class].double().cpu().numpy())
    masks = None
    if dropout > 0:
        masks = []
        for b in range(ntasks):
            sup = [(mask_array(seed, b, s, 0, NK, 256, dropout), mask_array(seed, b, s, 1, NK, 64, dropout))
                   for s in range(steps)]
            masks.append(dict(sup=sup, qry=(mask_array(seed, b, steps, 0, NQ, 256, dropout),
                                            mask_array(seed, b, steps, 1, NQ, 64, dropout))))
    params = {k: v.detach().double().cpu().numpy() for k, v in model.state_dict().items()}
    ref = episode_np.fumi_batch(params, fb, float(args.step_size), steps, tanh=bool(model.norm_hypernet), masks=masks,
                                want_grad=True, dtype=np.float64)
    rel = lambda x, y: float(np.abs(np.asarray(x, np.float64) - y).max() / max(np.abs(y).max(), 1e-30))
    gmax = max(np.abs(v).max() for v in ref["grads"].values())
    gerr = {}
    for k, p in model.named_parameters():
        g = p.grad.detach().double().cpu().numpy()
        gerr[k] = float(np.abs(g - ref["grads"][k]).max() / gmax) if k == "hyper_net.2.bias" else rel(g, ref["grads"][k])
    out = {"tasks": ntasks, "preds_equal": bool(np.array_equal(res["preds"].cpu().numpy(), ref["preds"])),
           "logits_relerr": rel(res["logits"].cpu().numpy(), ref["logits"]), "grad_relerr_max": max(gerr.values()),
           "grad_relerr": gerr, "checker": "oracle/episode_np.fumi_batch fp64, same dropout masks (fumi_b200/dropout.py)",
           "tolerance": {"logits": 1e-4, "grads": 2e-4}}
    out["ok"] = bool(out["preds_equal"] and out["logits_relerr"] < 1e-4 and out["grad_relerr_max"] < 2e-4)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="run this workload alone (default: fumi_train_5w5s as the headline + every other BASELINE config "
                         "under `secondary`)")
    ap.add_argument("--tasks", type=int, default=0, help="tasks per meta-batch per GPU (default: per workload)")
    ap.add_argument("--bank_images", type=int, default=195605)
    ap.add_argument("--bank_classes", type=int, default=673)
    ap.add_argument("--im_dim", type=int, default=2048)
    ap.add_argument("--text_dim", type=int, default=768)
    ap.add_argument("--dropout", type=float, default=0.25)
    ap.add_argument("--precision", type=int, default=int(os.environ.get("FUMI_PRECISION", "2")),
                    help="dense layers: 2 = tcgen05 with fp16 hi/lo bank planes (default), 1 = tcgen05 3xTF32, 0 = fp32 FMA")
    ap.add_argument("--cpu_budget_s", type=float, default=12.0)
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--no_kernel_pass", action="store_true")
    ap.add_argument("--no_secondary", action="store_true")
    ap.add_argument("--no_parity", action="store_true")
    ap.add_argument("--host_sampler", action="store_true",
                    help="e2e leg with the all-host native sampler instead of plan (host) + expand (device)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    solo = a.workload is not None
    if a.workload is None:
        a.workload = PRIMARY

    if a.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_leg(a.workload, a, a.steps, a.warmup, budget_s=240.0, min_s=30.0)
        metric, unit = metric_of(a.workload)
        w = WORKLOADS[a.workload]
        line = {"impl": "reference", "metric": metric, "value": r["value"], "unit": unit, "n_gpus": a.gpus,
                "steps": r["steps_done"], "warmup": a.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic iNat-Anim-shaped bank (fumi_b200.data.synth, RandomState(2022))",
                "config": {"workload": a.workload, "num_ways": w["N"], "num_shots": w["K"], "query_per_class": w["Q"],
                           "inner_steps": w["steps"], "tasks_per_batch_per_gpu": 4, "im_dim": a.im_dim,
                           "text_dim": a.text_dim, "hidden": [256, 64], "dropout": a.dropout if w["train"] else 0.0,
                           "parallelism": "reference: single process, host cores", "torch_threads": r["cores"],
                           "cpu": r["cpu"]},
                "cpu_baseline": {"value": r["value"], "unit": unit, "cores": r["cores"], "kind": r["kind"],
                                 "sample": r["sample"], "cpu": r["cpu"]},
                "e2e": {"value": r["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    ctx = Context(a)
    cpu_s = 0.0 if a.no_cpu_baseline else a.cpu_budget_s
    main_r = run_workload(ctx, a.workload, a.steps, a.warmup, not a.no_kernel_pass, cpu_s, not a.no_parity)
    secondary = {}
    if not solo and not a.no_secondary:
        for wl in SECONDARY:
            k = max(2, min(a.steps, 5 if WORKLOADS[wl]["steps"] >= 100 or wl == "fumi_train_20w5s" else a.steps))
            r = run_workload(ctx, wl, k, min(a.warmup, 3), not a.no_kernel_pass, min(cpu_s, 6.0), False)
            secondary[wl] = {key: r[key] for key in ("metric", "value", "unit", "ms_per_step", "steps", "warmup", "e2e",
                                                     "cpu_baseline", "gpu_launches", "config")}
            secondary[wl]["roofline"] = {"whole_step": {"achieved": r["roofline"]["achieved"], "frac": r["roofline"]["frac"]},
                                         "by_kernel": r["roofline"].get("by_kernel")}
            if "ranks_in_sync" in r:
                secondary[wl]["ranks_in_sync"] = r["ranks_in_sync"]
    strong = None
    if ctx.world > 1 and not solo and not a.no_secondary:
        # strong scaling: the single-GPU meta-batch (4096 tasks) split over the ranks, same all-reduce per step
        per = max(1, WORKLOADS[PRIMARY]["tasks"] // ctx.world)
        r = run_workload(ctx, PRIMARY, a.steps, a.warmup, False, 0.0, False, tasks_override=per)
        strong = {"scaling": "strong", "global_tasks_per_step": per * ctx.world, "tasks_per_gpu": per, "value": r["value"],
                  "unit": r["unit"], "ms_per_step": r["ms_per_step"], "e2e": r["e2e"], "ranks_in_sync": r.get("ranks_in_sync")}
    if ctx.rank == 0:
        line = {"metric": main_r["metric"], "value": main_r["value"], "unit": main_r["unit"], "n_gpus": ctx.world,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": main_r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic iNat-Anim-shaped bank (fumi_b200.data.synth, RandomState(2022)); random-init weights",
                "config": main_r["config"], "clocks": main_r["clocks"], "e2e": main_r["e2e"],
                "gpu_launches": main_r["gpu_launches"], "roofline": main_r["roofline"],
                "cpu_baseline": main_r["cpu_baseline"], "kernels": main_r["kernels"]}
        for key in ("parity", "ranks_in_sync"):
            if key in main_r:
                line[key] = main_r[key]
        if secondary:
            line["secondary"] = secondary
        if strong:
            line["strong_scaling"] = strong
        print(json.dumps(line))
    if ctx.world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    if ctx.rank == 0 and main_r.get("parity") is not None and not main_r["parity"]["ok"]:
        sys.exit(3)


if __name__ == "__main__":
    main()
