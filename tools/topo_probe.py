#!/usr/bin/env python
"""Host->device topology probe for the end-to-end scaling question (e2e 0.96 / 0.55 / 0.44 at
2 / 4 / 8 GPUs, aggregate H2D 98 / 112 / 181 GB/s).  Records what the box looks like (nvidia-smi topo -m, NUMA
nodes, the NUMA node of every GPU's PCI function) and measures pinned-host -> device bandwidth with one process
per GPU for: every GPU alone, GPUs {0..3}, {4..7}, {0,2,4,6}, all, and — when the box exposes more than one NUMA
node — with each process bound to its GPU's node vs to the other one.  One JSON document on stdout.

    python tools/topo_probe.py [--mb 512] [--reps 6]          # uses every visible GPU
"""
import argparse
import glob
import json
import multiprocessing as mp
import os
import subprocess
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=30).stdout
    except Exception as e:  # pragma: no cover
        return f"<{e}>"


def worker(gpu, bind, mb, reps, barrier, out_q):
    import torch
    from repurpose_b200 import affinity
    info = {"gpu": gpu, "numa_of_gpu": affinity.gpu_numa_node(gpu)}
    if bind == "own":
        info["bound"] = affinity.bind_to_gpu_numa(gpu)
    elif isinstance(bind, int):
        try:
            with open(f"/sys/devices/system/node/node{bind}/cpulist") as f:
                cpus = affinity._parse_cpulist(f.read()) & os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
            info["bound"] = {"numa_node": bind, "cpus": len(cpus)}
        except Exception as e:
            info["bound"] = {"error": str(e)}
    torch.cuda.set_device(gpu)
    n = mb * 1024 * 1024 // 4
    src = torch.empty(n, dtype=torch.float32).pin_memory()   # first touch under the affinity set above
    src.fill_(1.0)
    dst = torch.empty(n, dtype=torch.float32, device=f"cuda:{gpu}")
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    barrier.wait()
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier.wait()
    info["gbs"] = n * 4 * reps / dt / 1e9
    out_q.put(info)


def measure(gpus, bind, mb, reps):
    ctx = mp.get_context("spawn")
    barrier = ctx.Barrier(len(gpus))
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(g, bind if not isinstance(bind, dict) else bind[g], mb, reps, barrier, q))
             for g in gpus]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in procs), key=lambda r: r["gpu"])
    for p in procs:
        p.join()
    return {"gpus": gpus, "bind": bind if not isinstance(bind, dict) else "map", "aggregate_gbs": sum(r["gbs"] for r in res),
            "per_gpu_gbs": [round(r["gbs"], 1) for r in res], "numa_of_gpu": [r["numa_of_gpu"] for r in res],
            "bound": [r.get("bound") for r in res]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=512)
    ap.add_argument("--reps", type=int, default=6)
    a = ap.parse_args()
    import torch
    n = torch.cuda.device_count()
    nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
    doc = {"gpus": n, "numa_nodes": nodes, "cpus_allowed": len(os.sched_getaffinity(0)),
           "lscpu": [l for l in sh("lscpu").splitlines() if any(k in l for k in ("Model name", "Socket", "NUMA", "CPU(s):", "Thread"))],
           "topo": sh("nvidia-smi topo -m"),
           "pci_numa": {os.path.basename(os.path.dirname(p)): open(p).read().strip()
                        for p in glob.glob("/sys/bus/pci/devices/*/numa_node")
                        if os.path.exists(os.path.join(os.path.dirname(p), "class"))
                        and open(os.path.join(os.path.dirname(p), "class")).read().startswith("0x0302")},
           "meminfo": [l for l in sh("cat /proc/meminfo").splitlines()[:3]], "runs": []}
    sets = [[g] for g in range(n)]
    if n >= 4:
        sets += [list(range(4))]
    if n >= 8:
        sets += [list(range(4, 8)), [0, 2, 4, 6], [0, 1, 4, 5], list(range(8))]
    elif n > 1:
        sets += [list(range(n))]
    for gs in sets:
        doc["runs"].append(measure(gs, "own", a.mb, a.reps))
    if len(nodes) > 1 and n > 1:
        allg = list(range(n))
        doc["runs"].append(measure(allg, None, a.mb, a.reps))                 # unbound
        for node in nodes[:2]:
            doc["runs"].append(measure(allg, node, a.mb, a.reps))             # everybody on one node
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main()
