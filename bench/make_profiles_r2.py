#!/usr/bin/env python
"""Turn the raw outputs of the round-2 GPU calls (gpurun_out/r2c*.{json,jsonl,txt,csv,log}) into the tracked summaries under profiles/.
Re-runnable: every section is generated from whatever files exist."""
import csv
import glob
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def jl(path):
    out = []
    if os.path.exists(path):
        for l in open(path):
            l = l.strip()
            if l.startswith("{"):
                try:
                    out.append(json.loads(l))
                except ValueError:
                    pass
    return out


def bench(path):
    rows = jl(path)
    return rows[-1] if rows else None


def perf_rows(path):
    out = {}
    if not os.path.exists(path):
        return out
    for l in open(path):
        m = re.match(r"\s+(\d+)\s+\d+\s+\w+\s+(\w+)\s+\|\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)\s+(\d+)\s+\|\s+([\-\d.]+)\s+([\d.]+)\s+([\d.]+)\s+(\d+)", l)
        if m:
            out[int(m.group(1))] = {"algo": m.group(2), "oop_us": float(m.group(3)), "oop_busbw": float(m.group(5)), "ip_us": float(m.group(7)), "wrong": int(m.group(6)) + int(m.group(10))}
    return out


def size(n):
    for u, s in ((1 << 30, "GiB"), (1 << 20, "MiB"), (1 << 10, "KiB")):
        if n >= u:
            return f"{n / u:g} {s}"
    return f"{n} B"


def host_path():
    out = ["# End-to-end host path (`b200collAllReduceHost`): measurements", "",
           "Pinned host memory in, pinned host memory out, bf16 sum; device-timed with CUDA events on the caller's stream, max over ranks, 5 iterations after 2 warm-ups "
           "(`bench/e2e_hostpath.py`). `seq` = copy in, `b200collAllReduce`, copy back on one stream (what a user of a device-pointer API writes); `lib` = one call of "
           "`Comm.all_reduce_host`. `h2d` / `d2h` / `duplex` are the raw copy legs of the same buffers (duplex = both directions at once on two streams): the PCIe ceiling of the box.", ""]
    for tag, path, title in (("n1", "r2c2_host_default.jsonl", "1 x B200 (first thresholds: zero-copy <= 128 KiB, pipeline >= 2 MiB, chunk = total/8 in [1, 8] MiB)"),
                             ("n2", "r2c3_n2_host.jsonl", "2 x B200 (shipped thresholds: zero-copy <= 4 MiB where a Lamport kernel reaches, pipeline >= 8 MiB, chunk = total/6 in [4, 16] MiB)"),
                             ("n8", "r2c5_n8_host.jsonl", "8 x B200 (shipped thresholds)")):
        rows = [r for r in jl(os.path.join(G, path)) if "bytes" in r]
        if not rows:
            continue
        head = jl(os.path.join(G, path))[0]
        out += [f"## {title}", "", f"rank 0: GPU on NUMA node {head.get('numa_node_rank0')}, thread bound to {head.get('affinity')} CPUs (`{head.get('cpus_rank0')}`)", "",
                "| bytes | seq us | lib us | speed-up | lib algbw GB/s | h2d GB/s | d2h GB/s | duplex us |", "|---:|---:|---:|---:|---:|---:|---:|---:|"]
        for r in rows:
            out.append(f"| {size(r['bytes'])} | {r['seq_us']} | {r['lib_us']} | {r['speedup']} | {r['lib_algbw']} | {r.get('h2d_gbs', '')} | {r.get('d2h_gbs', '')} | {r.get('duplex_us', '')} |")
        out.append("")
    chunk = []
    for ck in (1024, 2048, 4096, 16384):
        rows = [r for r in jl(os.path.join(G, f"r2c2_host_c{ck}.jsonl")) if "bytes" in r]
        if rows:
            chunk.append((ck, rows))
    dflt = [r for r in jl(os.path.join(G, "r2c2_host_default.jsonl")) if r.get("bytes", 0) >= 16 << 20]
    if chunk:
        out += ["## Chunk size of the pipeline (1 x B200, `B200COLL_HOST_CHUNK_KB`)", "", "lib us per call; the sequential path takes 619 / 2520 / 9620 / 38430 us at these sizes. "
                "A chunk boundary costs about 20 us of bubble (event hand-over between the copy-in, compute and copy-back streams) and ~30 us of host enqueue time, "
                "so 1 MiB chunks give no gain at all (the host cannot enqueue 1024 chunks faster than they run) and the best chunk count is 4-8.", "",
                "| chunk | 16 MiB | 64 MiB | 256 MiB | 1 GiB |", "|---|---:|---:|---:|---:|"]
        for ck, rows in chunk:
            out.append(f"| {ck >> 10} MiB | " + " | ".join(str(r["lib_us"]) for r in rows) + " |")
        if dflt:
            out.append("| total/8 in [1, 8] MiB | " + " | ".join(str(r["lib_us"]) for r in dflt) + " |")
        out.append("")
    zc = [r for r in jl(os.path.join(G, "r2c2_host_zc512.jsonl")) if "bytes" in r]
    z0 = {r["bytes"]: r for r in jl(os.path.join(G, "r2c2_host_zc0.jsonl")) if "bytes" in r}
    if zc:
        out += ["## Zero-copy kernel vs copy engines for small messages (1 x B200)", "",
                "One kernel that reads the pinned input and writes the pinned output over PCIe (`lib`, zero-copy up to 512 KiB in this run) against copy + kernel + copy (`seq`, and `lib` with zero-copy off).", "",
                "| bytes | seq us | lib zero-copy us | lib with zero-copy off us |", "|---:|---:|---:|---:|"]
        for r in zc:
            if r["bytes"] <= 1 << 20:
                out.append(f"| {size(r['bytes'])} | {r['seq_us']} | {r['lib_us']} | {z0.get(r['bytes'], {}).get('lib_us', '')} |")
        out.append("")
    na = [r for r in jl(os.path.join(G, "r2c3_n2_host_noaff.jsonl")) if "bytes" in r]
    if na:
        out += ["## NUMA placement (2 x B200, both on one socket)", "", "`B200COLL_AFFINITY=0` and `torch.pin_memory()` buffers instead of `Comm.host_empty()`: "
                + ", ".join(f"{size(r['bytes'])}: {r['lib_us']} us" for r in na) + " (placed: see the 2-GPU table). With both GPUs on one socket the placement costs ~3 %; the 8-GPU table is where four unplaced ranks cross the socket link.", ""]
    open(os.path.join(P, "host_path.md"), "w").write("\n".join(out) + "\n")


def latency_ab():
    out = ["# A/B of the latency switches and of the copy-engine path", "",
           "`build/b200coll_perf --procs` (one process per GPU), bf16, out-of-place us per call, 200 timed iterations after 20 warm-ups for all-reduce, 20 after 5 for the data-movement ops. "
           "Run-to-run noise on these boxes is about +-0.5 us below 1 MiB and +-1 us above.", ""]
    for n in (2, 8):
        files = sorted(glob.glob(os.path.join(G, f"r2c*_n{n}_ab_ar_*.txt")))
        if not files:
            continue
        out += [f"## all-reduce, {n} x B200", ""]
        tabs = {re.search(r"ab_ar_(.*)_\.txt", f).group(1): perf_rows(f) for f in files}
        sizes = sorted(next(iter(tabs.values())))
        out += ["| setting | " + " | ".join(size(s) for s in sizes) + " |", "|---|" + "---:|" * len(sizes)]
        for k, t in tabs.items():
            out.append(f"| `{k.replace('_', ' ')}` | " + " | ".join(f"{t[s]['oop_us']:.2f}" if s in t else "" for s in sizes) + " |")
        algos = next(iter(tabs.values()))
        out += ["", "algorithm per size: " + ", ".join(f"{size(s)} {algos[s]['algo']}" for s in sizes), ""]
        for op in ("all_gather", "alltoall", "broadcast"):
            d, b = perf_rows(os.path.join(G, glob.glob(os.path.join(G, f"r2c*_n{n}_ab_{op}_default_.txt"))[0].split("/")[-1])) if glob.glob(os.path.join(G, f"r2c*_n{n}_ab_{op}_default_.txt")) else {}, {}
            fb = glob.glob(os.path.join(G, f"r2c*_n{n}_ab_{op}_B200COLL_BULK=0_.txt"))
            if fb:
                b = perf_rows(fb[0])
            if d and b:
                out += [f"### {op}, {n} x B200: copy-engine ring (`k_bulk`) vs LDG/STG push", "", "| bytes | k_bulk us | busbw GB/s | LDG/STG us | busbw GB/s | #wrong |", "|---:|---:|---:|---:|---:|---:|"]
                for s in sorted(d):
                    if s in b:
                        out.append(f"| {size(s)} | {d[s]['oop_us']:.2f} | {d[s]['oop_busbw']:.1f} | {b[s]['oop_us']:.2f} | {b[s]['oop_busbw']:.1f} | {d[s]['wrong'] + b[s]['wrong']} |")
                out.append("")
    open(os.path.join(P, "latency_ab.md"), "w").write("\n".join(out) + "\n")


def bench_tables():
    out = ["# bench.py, round 2: ours vs stock NCCL (defaults) vs stock NCCL on ncclMemAlloc + symmetric windows", "",
           "`python -m torch.distributed.run ... bench.py --gpus N --steps 10 --warmup 3 [--impl reference | reference-sym]`; bf16 sum all-reduce, out-of-place us (device-timed, max over ranks) "
           "and the end-to-end step (pinned host in, the whole result back in pinned host memory; ours: one `Comm.all_reduce_host` call, NCCL: copy, ncclAllReduce, copy).", ""]
    for n, pat in ((1, "r2c4_bench_{}.json"), (2, "r2c3_n2_bench_{}.json"), (4, "r2c5_n4_bench_{}.json"), (8, "r2c5_n8_bench_{}.json"), ("8 (final tree: device-side rendezvous before the timed region, --steps 20 --warmup 5)", "r2c7_n8_bench_{}.json")):
        arms = {a: bench(os.path.join(G, pat.format(a))) for a in ("ours", "reference", "reference-sym", "ref")}
        arms = {k: v for k, v in arms.items() if v}
        if "ref" in arms:
            arms["reference"] = arms.pop("ref")
        if "ours" not in arms or "reference" not in arms:
            continue
        o, r, s = arms["ours"], arms["reference"], arms.get("reference-sym")
        out += [f"## {n} x B200", "", f"| | ours | NCCL | NCCL sym |", "|---|---:|---:|---:|",
                f"| value (avg busbw GB/s) | {o['value']} | {r['value']} | {s['value'] if s else ''} |",
                f"| peak busbw GB/s | {o['peak_busbw']} | {r['peak_busbw']} | {s['peak_busbw'] if s else ''} |",
                f"| e2e avg busbw GB/s | {o['e2e']['value']} | {r['e2e']['value']} | {(s.get('e2e') or {}).get('value', 'not run') if s else ''} |",
                f"| verified (random bf16, 1 ulp of fp32 reference) | {o['verified_vs_torch_fp32']} | {r['verified_vs_torch_fp32']} | {s['verified_vs_torch_fp32'] if s else ''} |", "",
                "| bytes | algo | ours us | NCCL us | NCCL sym us | ours / best NCCL | ours e2e us | NCCL e2e us | e2e ratio |", "|---:|---|---:|---:|---:|---:|---:|---:|---:|"]
        st = {x["bytes"]: x for x in (s["table"] if s else [])}
        for a, b, c, d in zip(o["table"], r["table"], o["e2e"]["table"], r["e2e"]["table"]):
            best = min(b["oop_us"], st.get(a["bytes"], {}).get("oop_us", 1e30))
            out.append(f"| {size(a['bytes'])} | {a['algo']} | {a['oop_us']:.2f} | {b['oop_us']:.2f} | {st.get(a['bytes'], {}).get('oop_us', '')} | {best / a['oop_us']:.2f} | {c['e2e_us']:.1f} | {d['e2e_us']:.1f} | {d['e2e_us'] / c['e2e_us']:.2f} |")
        out.append("")
    open(os.path.join(P, "bench_r2.md"), "w").write("\n".join(out) + "\n")


def ncu_ranks():
    out = ["# One ncu per rank: what a profiler can and cannot see of a cross-GPU kernel", "",
           "`bench/ncu_ranks.sh`: every rank process of a one-rank-per-GPU run (`build/b200coll_perf` under RANK / WORLD_SIZE) is started under its OWN `ncu`, with a metric list that fits "
           "one pass, `--clock-control none --cache-control none`. Findings on this pool (driver 580.159, ncu 2025.2):", "",
           "* metric sets that fit one pass (DRAM bytes, L2 sectors arriving from the fabric, occupancy, launch shape) are collected per rank while the peers run un-profiled at full speed;",
           "* anything that needs a second pass fails with `==ERROR== UnknownError / Failed to profile` — kernel replay has to save and restore device memory, and a symmetric arena is "
           "VMM memory imported from other processes plus a multicast binding, which ncu cannot snapshot; the per-warp stall ratios, and `nvlrx__bytes` + `nvltx__bytes` + "
           "`gpu__time_duration` requested together, fell in this class on the 2-GPU runs (on ONE GPU the two NVLink counters alone are collected in a single pass, see the end: they exist, "
           "they just cannot be combined with a second counter domain here), so NVLink traffic is read from `lts__t_sectors_srcunit_ltcfabric` (32-byte sectors entering this GPU's L2 "
           "from the fabric: the responses to this GPU's peer loads, and peers' stores landing here);",
           "* the two ranks are not in lock-step under the tool (the profiled launch of one rank may meet a warm-up launch of the other), so only the rank whose capture shows the expected fabric volume is quoted.", ""]
    rows = []
    for f in sorted(glob.glob(os.path.join(G, "ncu_n*_mem_r*.csv"))):
        m = re.search(r"ncu_n(\d)_(.*)_mem_r(\d)\.csv", f)
        lines = [l for l in open(f) if l.startswith('"')]
        if not lines:
            continue
        per = {}
        for r in csv.DictReader(lines):
            per.setdefault((r["ID"], r["Kernel Name"].split("(")[0].replace("void ", "")[:70], r["Grid Size"], r["Block Size"]), {})[r["Metric Name"]] = r["Metric Value"]
        for k, v in per.items():
            try:
                rows.append((int(m.group(1)), m.group(2), int(m.group(3)), k[1], k[2], k[3], v.get("launch__registers_per_thread"), float(v["dram__bytes_read.sum"]), float(v["dram__bytes_write.sum"]),
                             float(v["lts__t_sectors_srcunit_ltcfabric.sum"]) * 32, v.get("sm__warps_active.avg.pct_of_peak_sustained_active")))
            except (KeyError, ValueError):
                pass
    if rows:
        out += ["| GPUs | run | rank | kernel | grid | block | regs | DRAM read MB | DRAM write MB | NVLink ingress MB (fabric sectors x 32) | warps active % |", "|---:|---|---:|---|---|---|---:|---:|---:|---:|---:|"]
        for r in rows:
            out.append(f"| {r[0]} | {r[1]} | {r[2]} | `{r[3]}` | {r[4]} | {r[5]} | {r[6]} | {r[7] / 1e6:.1f} | {r[8] / 1e6:.1f} | {r[9] / 1e6:.1f} | {r[10]} |")
        out += ["", "Reading the 2-GPU rows (64 MiB messages): `k_ar_twoshot` rank 0 receives 34.8 MB over NVLink for 33.6 MB algorithmic (its half of the buffer pulled from the peer; the pushed half is egress) — "
                "1.04 x; the P2P reduce-scatter pull shows the same 34.8 MB; the broadcast root's 32 MiB push arrives as 23.8 / 24.3 MB on the two ranks within the profiled window. `uncontrolled caches` means the DRAM columns "
                "include write-backs of lines the previous launch left dirty in L2; they bound, not equal, this launch's traffic. "
                "The 8-GPU rows are kept for the launch shapes and register counts only: with eight tools attached the ranks lose lock-step (a profiled launch meets peers that are seconds away, "
                "and since round 2 a kernel whose barrier times out returns without moving data), and switch-side traffic (`multimem.ld_reduce` / `multimem.st`) does not show up as fabric sectors at all — "
                "their fabric column reads ~0 although the same kernels move 843 GB/s when timed.", ""]
    probe = os.path.join(G, "r2c4_ncu_nvl_probe.txt")
    if os.path.exists(probe):
        txt = open(probe).read()
        out += ["## `nvlrx__bytes` / `nvltx__bytes` on one GPU", "", "```", *[l for l in txt.splitlines() if "nvl" in l or "ERROR" in l or "WARNING" in l][:12], "```", ""]
    open(os.path.join(P, "ncu_collectives.md"), "w").write("\n".join(out) + "\n")


def other_ops():
    out = ["# Every other collective at 8 x B200, both arms (round 2)", "",
           "`bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e --op all_gather --min 64K --extra-ops reduce_scatter,alltoall,broadcast,reduce,sendrecv,gather,scatter [--impl reference]`: "
           "bf16, 64 KiB ... 1 GiB x2, out-of-place and in-place where the op has one, bus bandwidth by the nccl-tests factors (rooted ops and sendrecv: busbw = algbw). "
           "`verified` = the op's result on pseudo-random data against a PyTorch fp32 reference (reductions within 1 bf16 ulp; NCCL's ring reductions round at every hop and miss that bound, "
           "which is what `False` on its arm means — see `bench_r2.md`).", ""]
    arms = {}
    for impl in ("ours", "reference"):
        try:
            ag = bench(os.path.join(G, f"r2c5_n8_ops_ag_{impl}.json"))
            ops = json.load(open(os.path.join(G, f"r2c5_n8_ops_{impl}.json")))["ops"]
            ops = {"all_gather": {"avg_busbw": ag["value"], "peak_busbw": ag["peak_busbw"], "verified": ag["verified_vs_torch_fp32"], "table": ag["table"]}, **ops}
            arms[impl] = ops
        except Exception:
            pass
    if len(arms) == 2:
        out += ["| op | ours avg / peak GB/s | NCCL avg / peak GB/s | ours / NCCL (avg) | verified ours / NCCL |", "|---|---:|---:|---:|---|"]
        for op in arms["ours"]:
            o, r = arms["ours"][op], arms["reference"].get(op)
            if r:
                out.append(f"| {op} | {o['avg_busbw']:.1f} / {o['peak_busbw']:.1f} | {r['avg_busbw']:.1f} / {r['peak_busbw']:.1f} | {o['avg_busbw'] / r['avg_busbw']:.2f} | {o['verified']} / {r['verified']} |")
        out.append("")
        for op in arms["ours"]:
            o, r = arms["ours"][op], arms["reference"].get(op)
            if not r:
                continue
            out += [f"## {op}", "", "| bytes | algo | ours us | busbw | NCCL us | busbw |", "|---:|---|---:|---:|---:|---:|"]
            rt = {x["bytes"]: x for x in r["table"]}
            for x in o["table"]:
                y = rt.get(x["bytes"])
                if y and x["bytes"] >= 1 << 20:
                    out.append(f"| {size(x['bytes'])} | {x['algo']} | {x['oop_us']:.1f} | {x['oop_busbw']:.1f} | {y['oop_us']:.1f} | {y['oop_busbw']:.1f} |")
            out.append("")
    open(os.path.join(P, "other_ops_n8_r2.md"), "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    host_path(); latency_ab(); bench_tables(); ncu_ranks(); other_ops()
    print("profiles written")
