"""ctypes binding of libjkq_b200.so (include/jkq.h).  No CPU fallback: a missing library or a
missing GPU raises."""

import ctypes as C
import os

from ._build import LIB, build_native, stale

_lib = None
ABI_VERSION = 5  # include/jkq.h JKQ_ABI_VERSION: bumped whenever a struct of the C ABI changes layout


class JkqConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("force_mode", C.c_int32), ("tolerance", C.c_double),
                ("node_capacity", C.c_uint32), ("value_capacity", C.c_uint32), ("item_capacity", C.c_uint32),
                ("cta_threads", C.c_uint32), ("cta_item_limit", C.c_uint32), ("gc_min", C.c_uint32),
                ("grid_blocks_per_sm", C.c_uint32), ("reserved", C.c_uint32)]


class JkqOp(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("target", C.c_uint32), ("ctrl_pos", C.c_uint64), ("ctrl_neg", C.c_uint64),
                ("m", C.c_double * 8), ("arg", C.c_uint32), ("reserved", C.c_uint32)]


class JkqStats(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in (
        "gates", "items", "nodes_created", "value_inserts", "gc_runs", "gc_nodes_copied", "retry_rounds",
        "max_frontier", "live_nodes", "live_values", "cta_launches", "grid_launches", "sample_launches")] + [
        ("apply_ms", C.c_double), ("sample_ms", C.c_double), ("sample_kernel_ms", C.c_double), ("norm", C.c_double),
        ("distinct_outcomes", C.c_uint64), ("thin_gates", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]


class JkqProgramInfo(C.Structure):
    _fields_ = [("n_qubits", C.c_uint32), ("n_clbits", C.c_uint32), ("n_ops", C.c_size_t), ("ops", C.POINTER(JkqOp)),
                ("n_qregs", C.c_uint32), ("n_cregs", C.c_uint32), ("n_snapshots", C.c_uint32),
                ("gate_after_measure", C.c_int32), ("first_measure", C.c_int64)]


class JkqSNode(C.Structure):
    _fields_ = [("c0", C.c_uint32), ("c1", C.c_uint32), ("q0", C.c_double)]


class JkqMultiStats(C.Structure):
    _fields_ = [("n_devices", C.c_uint32), ("merged", C.c_uint32), ("distinct_total", C.c_uint64), ("dd_bytes", C.c_uint64),
                ("broadcast_ms", C.c_double), ("sample_wall_ms", C.c_double), ("sample_kernel_ms_max", C.c_double),
                ("merge_ms", C.c_double), ("wall_ms", C.c_double)]


# every symbol include/jkq.h declares: (name, restype, argtypes)
_H = C.c_void_p
SYMBOLS = [
    ("jkq_abi_version", C.c_uint32, []),
    ("jkq_create", C.c_int, [C.POINTER(_H), C.POINTER(JkqConfig)]),
    ("jkq_destroy", None, [_H]),
    ("jkq_last_error", C.c_char_p, [_H]),
    ("jkq_load_circuit", C.c_int, [_H, C.c_uint32, C.POINTER(JkqOp), C.c_size_t]),
    ("jkq_set_option", C.c_int, [_H, C.c_char_p, C.c_int64]),
    ("jkq_apply", C.c_int, [_H]),
    ("jkq_sample", C.c_int, [_H, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int]),
    ("jkq_run", C.c_int, [_H, C.c_uint64, C.c_uint64]),
    ("jkq_get_counts", C.c_int, [_H, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_size_t, C.POINTER(C.c_size_t)]),
    ("jkq_release_histogram", C.c_int, [_H]),
    ("jkq_get_statevector", C.c_int, [_H, C.c_uint32, C.POINTER(C.c_double), C.c_size_t]),
    ("jkq_get_amplitudes", C.c_int, [_H, C.c_uint64, C.c_uint64, C.POINTER(C.c_double)]),
    ("jkq_get_snapshot_amplitudes", C.c_int, [_H, C.c_uint32, C.c_uint64, C.c_uint64, C.POINTER(C.c_double)]),
    ("jkq_get_snapshot_amplitudes_ordered", C.c_int, [_H, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint64, C.c_uint64, C.POINTER(C.c_double)]),
    ("jkq_get_level_counts", C.c_int, [_H, C.POINTER(C.c_uint32), C.c_size_t]),
    ("jkq_get_level_log", C.c_int, [_H, C.POINTER(C.c_uint32), C.c_size_t, C.POINTER(C.c_size_t)]),
    ("jkq_get_stats", C.c_int, [_H, C.POINTER(JkqStats)]),
    ("jkq_get_probabilities", C.c_int, [_H, C.c_uint32, C.POINTER(C.c_double), C.c_size_t, C.POINTER(C.c_size_t)]),
    ("jkq_get_probabilities_range", C.c_int, [_H, C.c_uint32, C.c_uint64, C.c_uint64, C.POINTER(C.c_double)]),
    ("jkq_get_qubit_probability", C.c_int, [_H, C.c_uint32, C.POINTER(C.c_double)]),
    ("jkq_get_mid_measurements", C.c_int, [_H, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    ("jkq_dd_export", C.c_int, [_H, C.POINTER(JkqSNode), C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    ("jkq_dd_import", C.c_int, [_H, C.c_uint32, C.POINTER(JkqSNode), C.c_uint32, C.c_uint32, C.POINTER(C.c_uint8)]),
    ("jkq_get_measure_map", C.c_int, [_H, C.POINTER(C.c_uint8), C.c_size_t]),
    ("jkq_dd_export_device", C.c_int, [_H, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    ("jkq_dd_import_device", C.c_int, [_H, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint8)]),
    ("jkq_sample_multi_gpu", C.c_int, [_H, C.POINTER(C.c_int32), C.c_uint32, C.c_uint64, C.c_uint64, C.POINTER(JkqMultiStats)]),
    ("jkq_counts_partition", C.c_int, [_H, C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)]),
    ("jkq_counts_merge", C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int]),
    ("jkq_batch_create", C.c_int, [C.POINTER(_H), C.c_int, C.c_int, C.POINTER(JkqConfig)]),
    ("jkq_batch_destroy", None, [_H]),
    ("jkq_batch_last_error", C.c_char_p, [_H]),
    ("jkq_batch_run", C.c_int, [_H, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(JkqOp), C.POINTER(C.c_uint64), C.c_uint64,
                                C.POINTER(C.c_uint64), C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32),
                                C.POINTER(C.c_int32), C.POINTER(C.c_uint64), C.POINTER(C.c_double)]),
    ("jkq_batch_run_shots", C.c_int, [_H, C.c_uint32, C.POINTER(JkqOp), C.c_size_t, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int,
                                      C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_int32), C.POINTER(C.c_uint64),
                                      C.POINTER(C.c_double)]),
    ("jkq_qasm_parse", C.c_int, [C.c_char_p, C.c_char_p, C.c_uint32, C.POINTER(_H)]),
    ("jkq_program_error", C.c_char_p, [_H]),
    ("jkq_program_free", None, [_H]),
    ("jkq_program_get_info", C.c_int, [_H, C.POINTER(JkqProgramInfo)]),
    ("jkq_program_get_register", C.c_int, [_H, C.c_int, C.c_uint32, C.c_char_p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    ("jkq_program_get_snapshot", C.c_int, [_H, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_size_t, C.POINTER(C.c_size_t)]),
    ("jkq_random_sector_bench", C.c_int, [_H, C.c_uint64, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
]


def load():
    """dlopen the in-tree library and bind every symbol; raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if stale():  # sources newer than the library (or no library yet): rebuild when nvcc is around
        try:
            from ._build import nvcc_path
            nvcc_path()
            have_nvcc = True
        except Exception:
            have_nvcc = False
        if have_nvcc:
            # (one builder at a time: the ranks of a torchrun launch all get here together)
            import fcntl
            with open(LIB + ".lock", "w") as lock:
                fcntl.flock(lock, fcntl.LOCK_EX)
                try:
                    if stale():
                        build_native()  # a failed rebuild must not leave Python bound to a stale binary with other struct layouts
                finally:
                    fcntl.flock(lock, fcntl.LOCK_UN)
    if not os.path.exists(LIB):
        raise ImportError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(the engine is CUDA only, there is no CPU fallback)" % LIB)
    lib = C.CDLL(LIB)
    lib.jkq_abi_version.restype = C.c_uint32
    got = lib.jkq_abi_version()
    if got != ABI_VERSION:
        raise ImportError("%s has ABI version %d, this package binds version %d: rebuild the library "
                          "(python -c 'import __graft_entry__ as g; g.build()')" % (LIB, got, ABI_VERSION))
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
