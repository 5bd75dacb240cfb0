"""Turn the ncu artefacts a gpurun call brought back into the committed summaries under profiles/.

    python tools/ncu_summary.py <round-tag> <launches.csv> <kv_len>=<report.ncu-rep> [...]

Writes profiles/<tag>_launches.md (per-kernel share of the step), profiles/<tag>_kernel_kv<kv>_raw.csv (ncu raw
page of the fused kernel) and profiles/ncu_summary.json (per-launch DRAM traffic, read by bench.py)."""
import collections, csv, json, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PROF = ROOT / "profiles"
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def launches(tag, path):
    lines = [l for l in open(path) if not l.startswith("==")]
    d = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1000 if row["Metric Unit"] == "ns" else v * 1000 if row["Metric Unit"] == "ms" else v
        d.setdefault(row["Kernel Name"], []).append(v)
    tot = sum(sum(v) for v in d.values())
    out = [f"# {tag}: every launch inside the timed region of `bench.py --steps 2` (value arm: 2 graph replays of 32 fused",
           "# launches; e2e arm: public-operator steps).  ncu --metrics gpu__time_duration.sum --clock-control none;",
           "# per-launch times are cold-cache and serialised: compare SHARES, not absolutes.", "",
           "| kernel | launches | mean us | share of GPU time |", "|---|---:|---:|---:|"]
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| `{k[:90]}` | {len(v)} | {sum(v)/len(v):.2f} | {100*sum(v)/tot:.1f}% |")
    (PROF / f"{tag}_launches.md").write_text("\n".join(out) + "\n")
    return d


def label_of(kv):
    return f"kv{kv}" if kv.isdigit() else kv        # "1024" -> kv1024 (headline MHA kernel); "gqa8k", "ffn" stay as given


def kernel(tag, kv, rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    (PROF / f"{tag}_kernel_{label_of(kv)}_raw.csv").write_text(raw)
    r = list(csv.reader(raw.splitlines()))
    hdr, units, rows = r[0], r[1], r[2:]
    res = {}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            vals = []
            for x in rows:
                try:
                    vals.append(float(x[i].replace(",", "")))
                except ValueError:
                    pass
            if vals:
                res[k] = {"unit": units[i], "mean": sum(vals) / len(vals), "per_launch": vals}
    return res


def main():
    tag = sys.argv[1]
    PROF.mkdir(exist_ok=True)
    launches(tag, sys.argv[2])
    summary = {}
    js = PROF / "ncu_summary.json"
    if js.exists():
        summary = json.loads(js.read_text())
    md = [f"# {tag}: ncu --set full --clock-control none, 3 launches each.  kv_len sections: `cfb::llama_decoder_layer_kernel<CHAT,4>`"
          " (Llama-2-7B); gqa*: `cfb::llama_decoder_layer_gqa2_kernel<SGLANG,4>` (Llama-3-8B); ffn: `cfb::llama_ffn_layer_kernel`", ""]
    for spec in sys.argv[3:]:
        kv, rep = spec.split("=")
        res = kernel(tag, kv, rep)
        mult = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}
        rd = res["dram__bytes_read.sum"]; wr = res["dram__bytes_write.sum"]
        traffic = rd["mean"] * mult[rd["unit"]] + wr["mean"] * mult[wr["unit"]]
        summary[f"traffic_bytes_{label_of(kv)}"] = int(traffic)
        summary[f"dram_read_bytes_{label_of(kv)}"] = int(rd["mean"] * mult[rd["unit"]])
        summary[f"source_{label_of(kv)}"] = f"profiles/{tag}_kernel_{label_of(kv)}_raw.csv"
        md += [f"## {'kv_len = ' + kv if kv.isdigit() else kv}", "", "| metric | unit | mean of 3 launches |", "|---|---|---:|"]
        for k, v in res.items():
            md.append(f"| {k} | {v['unit']} | {v['mean']:.3f} |")
        md.append("")
    (PROF / f"{tag}_kernel_summary.md").write_text("\n".join(md) + "\n")
    js.write_text(json.dumps(summary, indent=1) + "\n")
    print(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main()
