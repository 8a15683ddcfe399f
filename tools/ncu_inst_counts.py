"""Per-kernel-class instruction / DRAM counts of one bench step, from an ncu pass over bench.py itself.

    ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/inst.csv python bench.py --workload c4 --profile-frames 2
    python tools/ncu_inst_counts.py gpurun_out/inst.csv c4 2 RAYS_PER_STEP [existing.json] > profiles/r02_bench_inst_counts.json

bench.py divides these counts (thread instructions executed by a kernel class in one step) by the class's duration
that it measures live with CUDA events, which gives the lane-weighted issue utilisation of its roofline object.
Only the CULL=1 instantiations are counted (the frames bench.py times use the culled walk)."""
import collections
import csv
import json
import sys

CLASS_OF = (("rt_generate_kernel", "generate"), ("rt_walk_packet_kernel", "packet_walk"), ("rt_walk_kernel", "walk"),
            ("rt_longwalk_kernel", "long_walk"), ("rt_shade_kernel", "shade"), ("rt_resolve_kernel", "fold"))


def main():
    path, workload, frames, rays = sys.argv[1], sys.argv[2], int(sys.argv[3]), float(sys.argv[4])
    out = json.load(open(sys.argv[5])) if len(sys.argv) > 5 else {}
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    ix = {h: i for i, h in enumerate(rows[0])}
    per_launch = collections.OrderedDict()
    for r in rows[1:]:
        per_launch.setdefault(r[ix["ID"]], {"name": r[ix["Kernel Name"]]})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
    classes = {}
    for rec in per_launch.values():
        name = rec["name"].replace("void ", "")
        cls = next((c for k, c in CLASS_OF if name.startswith(k)), None)
        if cls is None:
            continue
        if cls != "fold" and "<(bool)1" not in name and "<true" not in name and "<1" not in name:
            continue                                        # exact-mode instantiations
        e = classes.setdefault(cls, {"thread_inst": 0.0, "warp_inst": 0.0, "dram_bytes": 0.0, "ncu_ms": 0.0, "launches": 0})
        e["thread_inst"] += rec.get("smsp__thread_inst_executed.sum", 0.0)
        e["warp_inst"] += rec.get("smsp__inst_executed.sum", 0.0)
        e["dram_bytes"] += rec.get("dram__bytes_read.sum", 0.0) + rec.get("dram__bytes_write.sum", 0.0)
        e["ncu_ms"] += rec.get("gpu__time_duration.sum", 0.0) / 1e6
        e["launches"] += 1
    for e in classes.values():
        for k in e:
            e[k] = e[k] / frames
    tot = sum(e["ncu_ms"] for e in classes.values()) or 1.0
    for e in classes.values():
        e["ncu_share"] = e["ncu_ms"] / tot
    out[workload] = {"classes": classes, "rays_per_step": rays, "frames_profiled": frames}
    out["source"] = "profiles/r02_bench_inst_counts.json: ncu over `python bench.py --workload W --profile-frames N` (tools/ncu_inst_counts.py); per step = sums / N"
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
