#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 QMDD engine (contract: see the task statement).

Workload (BASELINE.json configs[1]): QFT n=32 on the basis state |1010...>, cu1 expanded through
qelib1 -> 2528 ApplyGate calls, DD of 32 nodes: the apply-kernel latency path.
One step = one pass of the hot path = the whole circuit applied to |0...0> on the GPU.
metric = DD gate-applies per second (one apply = Simulator::ApplyGate, Simulator.cpp:466).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 (under torchrun): gate application does not shard (canonical DD), so every rank applies the
circuit on its own GPU (replicas, weak scaling); shot sampling DOES shard and is reported in `extra`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# The contract is ONE JSON line on stdout.  NCCL prints its version banner to the C-level stdout at any debug level
# from VERSION up, so file descriptor 1 is pointed at stderr for the whole run and the JSON line goes to a duplicate
# of the original stdout (emit()).
os.environ.setdefault("NCCL_DEBUG", "WARN")
sys.stdout.flush()
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(obj):
    _REAL_STDOUT.write(json.dumps(obj) + "\n")
    _REAL_STDOUT.flush()

METRIC = "dd_gate_applies_per_s"
UNIT = "gate-applies/s"
WORKLOAD = "QFT n=32 on |1010...>, 2528 U/CX applies (BASELINE configs[1])"
ALGO_BYTES_PER_ITEM = 320  # 10 random 32-byte sectors per processed frontier pair


def config_of(world):
    """The `config` object: identical in both arms (the driver compares them)."""
    return {"workload": WORKLOAD,
            "parallelism": "replicas x%d: gate application does not shard (one canonical DD); shots and circuit batches do -- see shots_scaling" % world,
            "l2": "256 MB buffer written between timed iterations (L2 flush); the reference arm runs on host cores"}


def workload():
    from qiskit_jku_provider_b200 import circuits as C
    return C.qft(32)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(circ, runs):
    """The UNMODIFIED reference engine (oracle/_ref/jku_ref, MPFR, single-threaded) on this box's host
    CPU; falls back to the FP64 C port when the reference binary is not there."""
    import oracle
    oracle.build()
    kind = "reference" if oracle.have_reference() else "port"
    fn = oracle.run_reference if kind == "reference" else oracle.run_oracle
    times = []
    for _ in range(runs):
        r = fn(circ)
        times.append(r["t_apply_s"])
    return kind, times


def executable_legs(device):
    """The process boundary the reference's Python backend crosses (qasm_simulator_jku.py:82-97): wall time of one
    `jku_simulator --simulate_qasm=FILE --shots=1024 --seed=1` process, ours (CUDA context creation included) and the
    unmodified reference executable (oracle/_ref/jku_simulator) on the same QASM text.  Best of 3 (reference: 1 run of
    the slow program)."""
    import subprocess
    import tempfile
    from qiskit_jku_provider_b200 import _build
    from qiskit_jku_provider_b200 import circuits as C
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "jku_simulator")
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        # the same executable behind its resident server (`jku_simulator --serve`): the short-lived process forwards
        # argv + program over a Unix socket, no CUDA start-up per invocation
        env = dict(os.environ, JKU_SIMULATOR_SOCKET=os.path.join(tmp, "jku.sock"))
        server = subprocess.Popen([_build.CLI, "--serve", "--idle=300", "--device=%d" % device], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        for _ in range(400):
            if os.path.exists(env["JKU_SIMULATOR_SOCKET"]):
                break
            time.sleep(0.05)
        for c, ref_runs in ((C.qft(32), 3), (C.random_circuit(18, 100, 1), 1)):
            path = os.path.join(tmp, c.name + ".qasm")
            with open(path, "w") as f:
                f.write(c.qasm())
            leg = {"applies": len(c)}
            for key, cmd, runs in (("b200", [_build.CLI, "--no_server", "--simulate_qasm=" + path, "--shots=1024", "--seed=1", "--device=%d" % device], 3),
                                   ("b200_resident_server", [_build.CLI, "--simulate_qasm=" + path, "--shots=1024", "--seed=1", "--device=%d" % device], 4),
                                   ("reference_cpu", [ref_exe, "--simulate_qasm=" + path, "--shots=1024", "--seed=1"], ref_runs)):
                if not os.path.exists(cmd[0]):
                    leg[key + "_wall_s"] = None
                    continue
                best = None
                for _ in range(runs):
                    t0 = time.perf_counter()
                    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900, env=env if key == "b200_resident_server" else None)
                    dt = time.perf_counter() - t0
                    if r.returncode == 0 and (best is None or dt < best):
                        best = dt
                leg[key + "_wall_s"] = best
            out[c.name] = leg
        server.terminate()
        try:
            server.wait(timeout=10)
        except Exception:
            server.kill()
    return out


def reference_per_shot(qasm, shots):
    """shots/s of the unmodified reference executable on a program it re-simulates per shot (needs qelib1.inc: the
    program is fed with the include replaced by the standard header text of the front end's built-in copy)"""
    import subprocess
    import tempfile
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "jku_simulator")
    if not os.path.exists(ref_exe):
        return None
    header = ("gate u3(theta,phi,lambda) q { U(theta,phi,lambda) q; }\ngate u2(phi,lambda) q { U(pi/2,phi,lambda) q; }\n"
              "gate u1(lambda) q { U(0,0,lambda) q; }\ngate cx c,t { CX c,t; }\ngate x a { u3(pi,0,pi) a; }\ngate z a { u1(pi) a; }\n"
              "gate h a { u2(0,pi) a; }\n")
    text = qasm.replace('include "qelib1.inc";\n', header)
    with tempfile.TemporaryDirectory() as tmp:
        t0 = time.perf_counter()
        r = subprocess.run([ref_exe, "--simulate_qasm", "--shots=%d" % shots, "--seed=1"], input=text, cwd=tmp, stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, universal_newlines=True, timeout=600)
        dt = time.perf_counter() - t0
    if r.returncode != 0:
        return None
    return {"shots": shots, "wall_s": dt, "shots_per_s": shots / dt}


def run_reference_arm(args):
    """bench.py --impl reference: the reference's own CPU engine on the host cores, same metric and config.
    The reference is single-threaded and gate application does not shard, so N GPUs are mirrored the way our
    own arm is: N independent replicas of the circuit, here N concurrent processes on N host cores (weak
    scaling), value = N x applies / slowest replica."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor
    import oracle
    oracle.build()
    circ = workload()
    kind = "reference" if oracle.have_reference() else "port"
    fn = oracle.run_reference if kind == "reference" else oracle.run_oracle
    replicas = max(1, args.gpus)

    def one_step():
        with ThreadPoolExecutor(replicas) as ex:
            return max(r["t_apply_s"] for r in ex.map(lambda _: fn(circ), range(replicas)))

    for _ in range(max(0, args.warmup)):
        one_step()
    t0 = time.perf_counter()
    times = [one_step() for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    t = sum(times) / len(times)
    value = replicas * len(circ) / t
    sample = "%s: full QFT-32 circuit (2528 applies) x %d steps x %d concurrent replica process(es), apply loop timed inside the engine (std::chrono), 1 thread each" % (
        "oracle/_ref/jku_ref = unmodified reference engine, MPFR 200-bit" if kind == "reference" else "oracle/qmdd_oracle.c FP64 port", args.steps, replicas)
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "mpfr200" if kind == "reference" else "f64", "data": "synthetic",
        "config": config_of(max(1, args.gpus)), "host_threads": replicas,
        "note": "the reference engine is single-threaded; one replica process per GPU of our arm",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": replicas, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": wall})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary measurements (ncu runs)")
    ap.add_argument("--cta-threads", type=int, default=0)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    from qiskit_jku_provider_b200 import circuits as C
    from qiskit_jku_provider_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    circ = workload()
    cfg = {"node_capacity": 1 << 20}
    if args.cta_threads:
        cfg["cta_threads"] = args.cta_threads
    eng = Engine(device=local, **cfg)
    eng.load(circ)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def step():
        flush.fill_(1)            # L2 flush between timed iterations (outside the event-timed region)
        torch.cuda.synchronize()
        eng.apply()               # CUDA-event timed inside the library, on the launching stream
        return eng.stats()

    for _ in range(max(3, args.warmup)):
        step()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    dev_ms, items, launches = [], 0, 0
    for _ in range(args.steps):
        before = eng.stats()
        st = step()
        dev_ms.append(st["apply_ms"])
        items = st["items"]
        launches += 1 + (st["cta_launches"] - before["cta_launches"]) + (st["grid_launches"] - before["grid_launches"])
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    my_ms = sum(dev_ms)
    if dist is not None:
        t = torch.tensor([my_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        my_ms = float(t[0])
    ms_per_step = my_ms / args.steps
    value = world * len(circ) / (ms_per_step * 1e-3)
    st = eng.stats()

    # ---- the state the LAST timed step left behind, against the oracle (outside the timed region) ----------
    parity = None
    if rank == 0:
        import oracle
        oracle.build()
        n_amp = 4096
        o = oracle.run_oracle(circ, amps=True, extra=["--amp-limit", str(n_amp), "--tie-tol"])
        oa = np.array([complex(*a) for a in o["amps"]])
        err = float(np.abs(eng.amplitudes(0, n_amp) - oa).max())
        lv = eng.level_counts()
        parity = {"checked": "the state left by the last timed step", "against": "oracle/qmdd_oracle.c (FP64 restatement, pinned to the reference)",
                  "amplitudes": n_amp, "max_amplitude_error": err, "tolerance": 1e-9, "level_counts_equal": lv == o["level_counts"],
                  "live_nodes_per_level": lv}
        if oracle.have_reference():
            r = oracle.run_reference(circ)
            parity["level_counts_equal_reference"] = lv == r["level_counts"]
        if not (err < 1e-9 and lv == o["level_counts"] and parity.get("level_counts_equal_reference", True)):
            raise SystemExit("bench.py: the timed run's result differs from the oracle: %s" % json.dumps(parity))

    # ---- end to end through the provider API: host Qobj in, host counts out ----------------------
    from qiskit_jku_provider_b200 import QasmSimulator
    from qiskit_jku_provider_b200.compat import ClassicalRegister, QuantumCircuit, QuantumRegister, assemble
    from qiskit_jku_provider_b200 import _native
    qr, cr = QuantumRegister(32, "q"), ClassicalRegister(32, "c")
    qc = QuantumCircuit(qr, cr, name="qft32")
    for i in range(32):
        if i % 4 in (0, 2):
            qc.x(qr[i])
    import math
    for i in range(32):
        qc.h(qr[i])
        for j in range(i + 1, 32):
            qc.cu1(math.pi / 2 ** (j - i), qr[j], qr[i])
    qc.measure(qr, cr)
    backend = QasmSimulator(silent=True, device=local, node_capacity=1 << 20)
    e2e_shots = 1024
    qobj = assemble(qc, shots=e2e_shots, seed=1)
    for _ in range(2):
        res = backend.run(qobj).result()
    barrier()
    copies0 = backend._get_engine().stats()
    t1 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(e2e_steps):
        res = backend.run(qobj).result()
    barrier()
    e2e_s = (time.perf_counter() - t1) / e2e_steps
    copies1 = backend._get_engine().stats()
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    e2e = {"value": world * len(circ) / e2e_s, "unit": UNIT,
           # counted inside the library: every cudaMemcpy of the handle adds its bytes (jkq_stats.h2d_bytes / d2h_bytes)
           "h2d_bytes_per_step": (copies1["h2d_bytes"] - copies0["h2d_bytes"]) // e2e_steps,
           "d2h_bytes_per_step": (copies1["d2h_bytes"] - copies0["d2h_bytes"]) // e2e_steps,
           "bytes_source": "measured: jkq_stats counters around the %d timed calls" % e2e_steps,
           "api": "QasmSimulator.run(qobj).result() -> get_counts, %d shots, host Qobj in / host counts out" % e2e_shots,
           "ms_per_call": 1e3 * e2e_s}

    # ---- secondary measurements ------------------------------------------------------------------
    extra = {}
    shots_scaling = {"n_gpus": world, "path": "torch.distributed: apply on rank 0 -> NCCL broadcast (device buffers) -> sample -> merge",
                     "note": "gate application is replicas-only (value); THIS block is what shards: compare shots_per_s_wall across --gpus N"}
    if not args.no_extra:
        from qiskit_jku_provider_b200 import multi_gpu
        bw = eng.random_sector_bench(1 << 32, 3) if rank == 0 else None
        if rank == 0:
            extra["random_sector"] = bw
        # ---- shot-sampling sweep (BASELINE configs[4]) sharded over the ranks: the numbers the scaling claim rests on.
        # wall = rank 0 applies the circuit -> NCCL broadcast of the sampling DD (device to device) -> every rank
        # samples its shot range -> histogram merge (dense all-reduce / owner exchange); kernel = the sampling kernel,
        # max over ranks.  1e9-shot QFT-32: ~0.9e9 distinct outcomes in a fused-slot HBM hash (34 GB at N = 1).
        for name, c2, counts_list in (("ghz24", C.ghz(24), (10 ** 6, 10 ** 8, 10 ** 9)), ("qft32", circ, (10 ** 6, 10 ** 8, 10 ** 9))):
            for shots in counts_list:
                best = None
                for attempt in range(3 if shots < 10 ** 9 else 2):
                    barrier()
                    t2 = time.perf_counter()
                    if dist is not None:
                        counts, sst = multi_gpu.sample_sharded(dist, eng, shots, 7, circuit=c2, sparse="owner")
                    else:
                        eng.load(c2)
                        eng.apply()
                        eng.sample(shots, seed=7)
                        sst = eng.stats()
                        counts = {"distinct_total": eng.distinct_outcomes()}
                    barrier()
                    w = time.perf_counter() - t2
                    k_ms = sst["sample_kernel_ms"]
                    if dist is not None:
                        t = torch.tensor([k_ms], dtype=torch.float64, device="cuda")
                        dist.all_reduce(t, op=dist.ReduceOp.MAX)
                        k_ms = float(t[0])
                    cur = {"shots": shots, "wall_s": w, "shots_per_s_wall": shots / w, "kernel_ms_max_over_ranks": k_ms,
                           "shots_per_s_kernel": shots / (k_ms * 1e-3)}
                    if isinstance(counts, dict) and "distinct_total" in counts:
                        cur["distinct_outcomes"] = counts["distinct_total"]
                    elif isinstance(counts, dict):
                        cur["distinct_outcomes"] = len(counts)
                    if best is None or cur["wall_s"] < best["wall_s"]:
                        best = cur
                shots_scaling["%s_%.0e" % (name, shots)] = best
        eng.free_histogram()
        if rank == 0:
            # the CPU side of the same sweep: MeasureAll of the unmodified reference (Simulator.cpp:148-224), one
            # thread, 2e4 shots measured; larger counts are this rate extrapolated (labelled so)
            import oracle as _o
            if _o.have_reference():
                for name, c2 in (("ghz24", C.ghz(24)), ("qft32", circ)):
                    r = _o.run_reference(c2, shots=20000, seed=1)
                    shots_scaling["cpu_reference_" + name] = {"shots_measured": 20000, "t_sample_s": r["t_sample_s"],
                                                              "shots_per_s": 20000 / r["t_sample_s"], "cores": 1,
                                                              "note": "1e6..1e9 shots: extrapolated from this rate"}
        # batch of independent circuits (BASELINE configs[4]): sharded over the SMs of each GPU and over the GPUs
        from qiskit_jku_provider_b200.batch import BatchRunner
        from qiskit_jku_provider_b200.engine import circuit_ops
        n_batch = 1024
        batch = [C.random_circuit(18, 40, s) for s in range(n_batch)]
        for c2 in batch:
            circuit_ops(c2)
        with BatchRunner(workers=148, device=local) as br:
            br.run(batch, shots=1024, seed=1, rank=rank, world=world)  # warm-up: sizes every buffer
            barrier()
            t2 = time.perf_counter()
            _, b_gates, b_ms = br.run(batch, shots=1024, seed=1, rank=rank, world=world)
            barrier()
            b_wall = time.perf_counter() - t2
            fb = br.fallbacks
            b_kernel_only, b_fallback = br.last_kernel_ms, br.last_fallback_ms
        # the same batch through the provider API: ONE Qobj of 1 024 experiments into QasmSimulator(statevector=False)
        # .run(qobj).result() -- named gates in, counts out (rank 0, its own GPU; `devices=` spreads it inside one process)
        api = None
        if rank == 0:
            import random as _random
            from qiskit_jku_provider_b200 import QasmSimulator as _QS
            from qiskit_jku_provider_b200.compat import ClassicalRegister as _CR, QuantumCircuit as _QC, QuantumRegister as _QR, assemble as _asm
            names = [k for k in sorted(C.RANDOM_OPS) if C.RANDOM_OPS[k][0] <= 18]
            qcs = []
            for s_ in range(n_batch):  # the draws of circuits.random_circuit(18, 40, s_), kept as named instructions
                rng = _random.Random(s_)
                qr_, cr_ = _QR(18, "q"), _CR(18, "c")
                qc_ = _QC(qr_, cr_, name="random18_40_s%d" % s_)
                for _ in range(40):
                    nm = rng.choice(names)
                    nregs, nparams = C.RANDOM_OPS[nm]
                    pr = [rng.random() for _ in range(nparams)]
                    qc_.data.append((nm, pr, rng.sample(range(18), nregs), []))
                qc_.measure(qr_, cr_)
                qcs.append(qc_)
            qobj_b = _asm(qcs, shots=1024, seed=1)
            be = _QS(silent=True, device=local, statevector=False, node_capacity=1 << 20)
            be.run(qobj_b).result()  # warm-up: engines, buffers
            t3 = time.perf_counter()
            res_b = be.run(qobj_b).result()
            api_wall = time.perf_counter() - t3
            # the same seed for every experiment (the backend's rule), so compare against the one-at-a-time engine on a few
            ok = True
            for s_ in (0, 17, 170, 1023):
                eng.load(batch[s_])
                eng.run(1024, 1)
                ok = ok and res_b.to_dict()["results"][s_]["data"]["counts"] == {hex(k_): v_ for k_, v_ in eng.counts().items()}
            api = {"call": "QasmSimulator(statevector=False).run(qobj of 1024 experiments).result()", "wall_s": api_wall,
                   "circuits_per_s": n_batch / api_wall, "gate_applies_per_s": sum(len(c2) for c2 in batch) / api_wall,
                   "counts_equal_one_at_a_time_engine": bool(ok), "metadata": res_b.to_dict()["results"][0]["metadata"]["batch"]}
        tt = torch.tensor([float(b_gates), b_ms], dtype=torch.float64, device="cuda")
        if dist is not None:
            g_all = tt.clone()
            dist.all_reduce(g_all, op=dist.ReduceOp.SUM)
            m_all = tt.clone()
            dist.all_reduce(m_all, op=dist.ReduceOp.MAX)
            b_gates, b_ms = float(g_all[0]), float(m_all[1])
        extra["batch_1024_random18_40"] = {
            "circuits": n_batch, "shots_each": 1024, "gate_applies": int(b_gates), "kernel_ms_max_over_ranks": b_ms,
            "gate_applies_per_s_kernel": b_gates / (b_ms * 1e-3), "wall_s_host_ops_in_counts_out": b_wall,
            "gate_applies_per_s_wall": b_gates / b_wall, "circuits_per_s_wall": n_batch / b_wall, "engines_per_gpu": 148,
            "full_engine_fallbacks_rank0": fb,
            "k_batch_ms_rank0": b_kernel_only, "fallback_apply_ms_rank0": b_fallback, "through_the_provider_api": api,
            "note": "one persistent kernel per GPU, one CTA per SM pulling circuits from a queue (longest gate list first); circuits split "
                    "i % world == rank; kernel_ms = k_batch + the circuits that outgrew one SM, redone from gate 0 on up to 4 full engines side by side (that part by the host clock: an upper bound)"}
        if rank == 0:
            # the CPU side of the batch: the unmodified reference, one process per circuit like its Python
            # backend runs them, P processes side by side on the host cores; a bounded sample of the same batch
            import oracle as _o
            from concurrent.futures import ThreadPoolExecutor
            if _o.have_reference():
                sample_n = 64
                for procs in sorted({1, min(16, os.cpu_count() or 1)}):
                    t2 = time.perf_counter()
                    with ThreadPoolExecutor(procs) as ex:
                        list(ex.map(lambda c4: _o.run_reference(c4, shots=1024, seed=1), batch[:sample_n]))
                    dt = time.perf_counter() - t2
                    extra["batch_1024_random18_40"]["cpu_reference_%d_processes" % procs] = {
                        "circuits_sampled": sample_n, "wall_s": dt, "circuits_per_s": sample_n / dt,
                        "gate_applies_per_s": sum(len(c4) for c4 in batch[:sample_n]) / dt,
                        "note": "oracle/_ref/jku_ref (unmodified reference), apply + 1024 shots per circuit, process start included"}
            for c3 in (C.ghz(24), C.grover(20, 20), C.random_circuit(18, 100, 1), C.random_circuit(18, 120, 1), C.random_circuit(18, 160, 1)):
                eng.load(c3)
                eng.apply()
                eng.apply()
                s3 = eng.stats()
                extra["apply_" + c3.name] = {"applies": len(c3), "apply_ms": s3["apply_ms"], "us_per_gate": 1e3 * s3["apply_ms"] / len(c3),
                                             "gate_applies_per_s": len(c3) / (s3["apply_ms"] * 1e-3), "items": s3["items"],
                                             "thin_gates": s3["thin_gates"], "nodes_created": s3["nodes_created"],
                                             "nodes_created_per_s": s3["nodes_created"] / (s3["apply_ms"] * 1e-3),
                                             "cta_launches_total": s3["cta_launches"], "grid_launches_total": s3["grid_launches"]}
            # the one config on which the random-sector roofline binds: random n=18, 160 ops, grid path
            g3 = extra["apply_random18_160_s1"]
            sect = extra["random_sector"]["random_sector_gbs"]
            ach = g3["items"] * ALGO_BYTES_PER_ITEM / (g3["apply_ms"] * 1e-3) / 1e9
            extra["roofline_grid"] = {"workload": "random n=18, 160 ops, seed 1 (107 k live nodes, 2 M frontier pairs): k_run_gates<MODE_GRID>",
                                      "bound": "hbm (random 32-byte sectors)", "achieved": ach, "peak": sect, "unit": "GB/s", "frac": ach / sect,
                                      "algorithmic_bytes": "%d frontier items x %d B" % (g3["items"], ALGO_BYTES_PER_ITEM),
                                      "peak_source": "jkq_random_sector_bench on this GPU, this run",
                                      "traffic": "profiles/ (ncu --set full of this kernel)"}
            # ... and the same kernel where the frontiers are wide enough to occupy the grid: random n=22, 200 ops (47 M frontier
            # pairs, 42 M nodes created; its own engine: 2^25 nodes, 2^26 reals)
            try:
                from qiskit_jku_provider_b200.engine import Engine as _Engine
                c5 = C.random_circuit(22, 200, 1)
                with _Engine(device=local, node_capacity=1 << 25, value_capacity=1 << 26) as big:
                    big.load(c5)
                    big.apply()
                    s5 = big.stats()
                ach5 = s5["items"] * ALGO_BYTES_PER_ITEM / (s5["apply_ms"] * 1e-3) / 1e9
                extra["roofline_grid_wide"] = {"workload": "random n=22, 200 ops, seed 1: k_run_gates<MODE_GRID>", "applies": len(c5),
                                               "apply_ms": s5["apply_ms"], "items": s5["items"], "items_per_s": s5["items"] / (s5["apply_ms"] * 1e-3),
                                               "nodes_created": s5["nodes_created"], "nodes_created_per_s": s5["nodes_created"] / (s5["apply_ms"] * 1e-3),
                                               "gc_runs": s5["gc_runs"], "bound": "hbm (random 32-byte sectors)", "achieved": ach5, "peak": sect,
                                               "unit": "GB/s", "frac": ach5 / sect,
                                               "algorithmic_bytes": "%d frontier items x %d B" % (s5["items"], ALGO_BYTES_PER_ITEM),
                                               "peak_source": "jkq_random_sector_bench on this GPU, this run"}
            except Exception as e:  # (a secondary figure: never the reason for a bench without its headline)
                extra["roofline_grid_wide"] = {"error": str(e)[:200]}

        # per-shot re-simulation of a circuit with gates behind measurements: every shot is an
        # instance of the circuit on its own SM (k_batch_shots), shots split over the ranks
        from qiskit_jku_provider_b200.qasm import parse_qasm
        prog = parse_qasm(C.teleport_qasm())
        ps_shots = 200000
        with BatchRunner(workers=1 << 16, device=local) as br:
            multi_gpu.run_shots_sharded(dist, br, prog.n_qubits, prog.ops, 2000, 3)
            barrier()
            t2 = time.perf_counter()
            merged, failed, ps_ms = multi_gpu.run_shots_sharded(dist, br, prog.n_qubits, prog.ops, ps_shots, 3)
            barrier()
            ps_wall = time.perf_counter() - t2
        if dist is not None:
            tt = torch.tensor([ps_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ps_ms = float(tt[0])
        extra["per_shot_teleport"] = {"shots": ps_shots, "kernel_ms_max_over_ranks": ps_ms, "shots_per_s_kernel": ps_shots / (ps_ms * 1e-3),
                                      "wall_s_ops_in_counts_out": ps_wall, "shots_per_s_wall": ps_shots / ps_wall,
                                      "q2_always_one": all((k >> 2) & 1 for k in merged), "shots_left_to_full_engine": len(failed),
                                      "note": "3 mid-circuit measurements + 2 conditioned gates per shot, handled on the device"}
        if rank == 0:
            extra["per_shot_teleport"]["reference_cpu"] = reference_per_shot(C.teleport_qasm(), 400)
            extra["executable"] = executable_legs(local)

    if rank == 0:
        peaks, peaks_src = measured_peaks()
        achieved = items * ALGO_BYTES_PER_ITEM / (ms_per_step * 1e-3) / 1e9
        traffic = None
        try:  # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu --set full capture
            with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
                t = json.load(f)["k_run_gates<MODE_CTA>"]
            traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
        except Exception:
            pass
        roof = {"bound": "latency", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peaks_src,
                "us_per_gate": 1e3 * ms_per_step / len(circ),
                "kernel": "k_run_gates<MODE_CTA> (persistent single-CTA circuit kernel, one warp: 1 launch per step)",
                "algorithmic_bytes": "%d frontier items/step x %d B" % (items, ALGO_BYTES_PER_ITEM),
                "note": "NOT HBM-bound: the DD (32 nodes) and every table live in one SM's shared memory, HBM sees %s bytes per "
                        "circuit, so achieved / peak is reported for the contract only.  One warp runs a dependent chain per gate; "
                        "the figure of merit is us_per_gate (reference 28 us, FP64 CPU port 10 us).  The config where the "
                        "random-sector roofline binds is in extra.roofline_grid" % (traffic if traffic is not None else "~0.5 M")}
        if "random_sector" in extra:
            roof["random_sector_gbs_measured"] = extra["random_sector"]["random_sector_gbs"]
            roof["frac_of_random_sector"] = achieved / extra["random_sector"]["random_sector_gbs"]
        cb_kind, cb_times = cpu_reference_run(circ, 5)
        cb_t = min(cb_times)
        cpu = {"value": len(circ) / cb_t, "unit": UNIT, "cores": 1, "kind": cb_kind,
               "sample": "full QFT-32 circuit, best of 5 runs of %s, apply loop only, 1 thread (engine is single-threaded)" % (
                   "oracle/_ref/jku_ref (unmodified reference, MPFR)" if cb_kind == "reference" else "oracle/qmdd_oracle.c"),
               "us_per_gate": 1e6 * cb_t / len(circ)}
        # the same algorithm in FP64 on one host core (oracle/qmdd_oracle.c): what the GPU has to beat once the
        # MPFR -> FP64 change is taken out of the comparison
        import oracle as _oracle
        pt = min(_oracle.run_oracle(circ)["t_apply_s"] for _ in range(5))
        cpu_port = {"value": len(circ) / pt, "unit": UNIT, "cores": 1, "kind": "port", "us_per_gate": 1e6 * pt / len(circ),
                    "sample": "full QFT-32 circuit, best of 5 runs of oracle/qmdd_oracle.c (FP64 restatement of the reference), apply loop only, 1 thread"}
        for nm, c3 in (("grover20_20", C.grover(20, 20)), ("random18_120_s1", C.random_circuit(18, 120, 1))):
            pt3 = min(_oracle.run_oracle(c3)["t_apply_s"] for _ in range(2))
            cpu_port["other_" + nm] = {"applies": len(c3), "us_per_gate": 1e6 * pt3 / len(c3), "gate_applies_per_s": len(c3) / pt3}
        out = {"shots_scaling": shots_scaling, "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
               "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic",
               "config": config_of(world), "us_per_gate": 1e3 * ms_per_step / len(circ),
               "timing": "CUDA events on the library's launch stream around init + circuit kernels, summed over steps, max over ranks",
               "wall_s_timed_region": wall,
               "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu,
               "cpu_baseline_port": cpu_port, "parity": parity,
               "engine_stats": {k: st[k] for k in ("gates", "items", "nodes_created", "value_inserts", "gc_runs", "retry_rounds",
                                                   "max_frontier", "live_nodes")},
               "extra": extra}
        emit(out)
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
