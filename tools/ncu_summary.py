"""Summarise an `ncu --page raw --csv` dump into profiles/*.json (per-launch duration, DRAM traffic, pipe activity).
   python tools/ncu_summary.py raw.csv out.json "<source description>" workload tasks"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try:
        return float(r[idx[k]].replace(",", ""))
    except (KeyError, ValueError):
        return None
names = {"episode_fwd_v2": "fumi_episode_fwd", "episode_bwd_v2": "fumi_episode_bwd", "gram_tc_kernel": "fumi_gram", "gram_kernel": "fumi_gram", "episode_fwd_mma16": "fumi_episode_fwd", "episode_fwd_f16": "fumi_episode_fwd",
         "episode_bwd_mma16": "fumi_episode_bwd", "gemm_x3_kernel<1>": "fumi_gemm_f16x3",
         "gemm_x3_kernel<0>": "fumi_gemm_tf32x3", "gemm_tf32x3": "fumi_gemm_tf32x3", "sampler_expand": "fumi_sampler_expand"}
out = {"source": sys.argv[3], "workload": sys.argv[4], "tasks": int(sys.argv[5]), "kernels": {}}
bscale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
tscale = {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}
for r in rows[2:]:
    kn = r[idx["Kernel Name"]]
    key = next((v for k, v in names.items() if k in kn), kn[:40])
    rec = {"kernel_name": kn[:80],
           "duration_ms": f(r, "gpu__time_duration.sum") * tscale[units[idx["gpu__time_duration.sum"]]],
           "dram_read_bytes": f(r, "dram__bytes_read.sum") * bscale[units[idx["dram__bytes_read.sum"]]],
           "dram_write_bytes": f(r, "dram__bytes_write.sum") * bscale[units[idx["dram__bytes_write.sum"]]],
           "dram_throughput_pct": f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
           "tensor_pipe_active_pct": f(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
           "issue_active_pct": f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
           "smem_wavefronts": f(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
           "smem_bank_conflicts": f(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
           "l2_hit_pct": f(r, "lts__t_sector_hit_rate.pct"),
           "regs_per_thread": f(r, "launch__registers_per_thread"),
           "dyn_smem_kb": f(r, "launch__shared_mem_per_block_dynamic")}
    rec["traffic_bytes"] = rec["dram_read_bytes"] + rec["dram_write_bytes"]
    out["kernels"].setdefault(key, []).append(rec)
json.dump(out, open(sys.argv[2], "w"), indent=1)
for k, v in out["kernels"].items():
    for x in v:
        print(f"{k:22s} {x['duration_ms']:7.3f} ms  dram {x['traffic_bytes'] / 1e9:6.3f} GB ({x['dram_throughput_pct']}%)  tensor {x['tensor_pipe_active_pct']}%  issue {x['issue_active_pct']}%")
