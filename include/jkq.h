/* jkq.h -- C ABI of the B200-native QMDD engine (libjkq_b200.so).
 *
 * This is the boundary the reference crosses with a subprocess pipe:
 *   qiskit_jku_provider/qasm_simulator_jku.py:82-97  JKUSimulatorWrapper.run()  argv + QASM on stdin,
 *                                                   JSON on stdout
 *   src/main.cpp:75-150                              srand, QMDDinit, QASMsimulator, Simulate(shots)
 * Each entry point names the reference routine it replaces.  Plain pointers and sizes only; the
 * caller owns every output buffer; a handle owns all device memory and is bound to one CUDA device.
 * A handle is not thread-safe; distinct handles may be used from distinct threads.
 * Every function returns 0 on success or a JKQ_E_* code; jkq_last_error() gives the text.
 * There is no CPU path: jkq_create fails when no CUDA device is usable.
 */
#ifndef JKQ_H
#define JKQ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct jkq_handle jkq_handle;

enum {
  JKQ_E_OK = 0,
  JKQ_E_NODE_FULL = 1,   /* node arena exhausted even after a collection */
  JKQ_E_UTAB_FULL = 2,
  JKQ_E_VALUE_FULL = 3,  /* complex-value table exhausted */
  JKQ_E_ITEM_FULL = 4,   /* frontier pool exhausted */
  JKQ_E_DTAB_FULL = 5,
  JKQ_E_PEND_FULL = 6,
  JKQ_E_BARRIER_TIMEOUT = 7,
  JKQ_E_ZERO_STATE = 8,  /* "numerical instabilities led to a 0-vector", Simulator.cpp:158-161 */
  JKQ_E_INTERNAL = 9,
  JKQ_E_RETRY_LIMIT = 10, /* watchdog: one level kept producing new values for 512 rounds */
  JKQ_E_VALUE_RANGE = 11, /* a weight of magnitude >= 1e8 or NaN: numerical breakdown of the state, not a capacity problem */
  JKQ_E_CUDA = 100,      /* CUDA runtime error or no device */
  JKQ_E_ARGUMENT = 101,
  JKQ_E_STATE = 102,     /* call made in the wrong order */
  JKQ_E_CAPACITY = 103   /* caller buffer too small */
};

/* One configuration struct at handle creation replaces the reference's compile-time limits and CLI
 * flags (QMDDpackage.h:35-57 MAXN/NBUCKET/GCLIMIT1, QMDDcomplexD.cpp:336 Ctol, main.cpp:108-113
 * --precision).  Zero in any field selects the default. */
typedef struct {
  int32_t device;           /* CUDA ordinal                                   default 0       */
  int32_t force_mode;       /* 0 auto, 1 single-CTA path only, 2 grid path only               */
  double tolerance;         /* complex-table tolerance Ctol                   default 1e-10   */
  uint32_t node_capacity;   /* nodes per arena (power of two)                 default 1<<22   */
  uint32_t value_capacity;  /* distinct reals (power of two)                  default 1<<20   */
  uint32_t item_capacity;   /* frontier work items (power of two)             default = nodes */
  uint32_t cta_threads;     /* threads of the wide single-CTA launch          default 256     */
  uint32_t cta_item_limit;  /* frontier size that escalates to the grid path  default 4096    */
  uint32_t gc_min;          /* never collect below this many allocated nodes  default 1<<15   */
  uint32_t grid_blocks_per_sm; /* grid path CTAs per SM                       default 2       */
  uint32_t reserved;        /* 1: always launch the single-CTA path with cta_threads (default: one warp until a
                               level of a gate outgrows it, then cta_threads)                              */
} jkq_config;

enum {
  JKQ_OP_GATE = 0, JKQ_OP_SNAPSHOT = 1, JKQ_OP_MEASURE = 2, JKQ_OP_BARRIER = 3, JKQ_OP_MEASURE_ONE = 4, JKQ_OP_RESET = 5,
  JKQ_OP_IF = 6, JKQ_OP_PROBABILITIES = 7
};

/* One operation of a circuit.
 *  GATE      2x2 matrix m (row major, re/im pairs) on `target`, applied where every qubit in ctrl_pos
 *            is 1 and every qubit in ctrl_neg is 0 -- the general form QMDDmvlgate builds
 *            (QMDDpackage.cpp:1746-1828); U and CX of the reference front end
 *            (QASMsimulator.cpp:281-357) are the cases {no control} and {m = X, one positive control}.
 *  SNAPSHOT  record the full state vector in slot `arg` (QASMsimulator.cpp:973-1037).
 *  MEASURE   qubit `target` -> classical bit `arg`; counts are sampled from the final state
 *            (QASMsimulator.cpp:487-500).
 *  BARRIER   no-op (QASMsimulator.cpp:944-949).
 *  MEASURE_ONE  measure qubit `target` NOW and collapse the state (Simulator::MeasureOne, Simulator.cpp:226-289):
 *            outcome drawn with P(0) = AssignProbsOne from the Philox stream (option "measure_seed", op index),
 *            state projected and rescaled by 1/sqrt(p); the outcome goes to classical bit `arg`
 *            (jkq_get_mid_measurements).
 *  RESET     Simulator::ResetQubit (Simulator.cpp:291-339): the reference's reset is a PROJECTION, not a measurement --
 *            the state is projected on |0> of `target` and rescaled by 1/sqrt(P(0)); only when P(0) = 0 an X comes
 *            first.  No random number is drawn.  Option "physical_reset" = 1 selects measure-then-X instead.
 *  IF        `if (creg == k) qop;` (QASMsimulator.cpp:952-971): the next `arg` ops (0 reads as 1; GATE, MEASURE_ONE,
 *            RESET or BARRIER) are executed only when (classical bits & ctrl_pos) == ctrl_neg, the classical bits
 *            being the MEASURE_ONE outcomes so far (unwritten bits read 0).  The condition is evaluated once.
 *  PROBABILITIES  record, in slot `arg`, the probability of every assignment of the qubits in the mask ctrl_pos
 *            (Simulator::GetProbability per assignment, QASMsimulator.cpp:990-1011): 2^k doubles, bit j of the index =
 *            the j-th lowest qubit of the mask (jkq_get_probabilities). */
typedef struct {
  uint32_t kind;
  uint32_t target;
  uint64_t ctrl_pos;
  uint64_t ctrl_neg;
  double m[8];
  uint32_t arg;
  uint32_t reserved;
} jkq_op;

typedef struct {
  uint64_t gates;            /* ApplyGate calls (Simulator.cpp:466)                            */
  uint64_t items;            /* frontier pairs processed (the reference's CTlook, :993)        */
  uint64_t nodes_created;    /* unique-table inserts (UTlookups - UTmatch, :708,730)           */
  uint64_t value_inserts;    /* reals appended to the value table                              */
  uint64_t gc_runs;
  uint64_t gc_nodes_copied;
  uint64_t retry_rounds;     /* serial value-insert sections                                   */
  uint64_t max_frontier;     /* largest per-gate item count                                    */
  uint64_t live_nodes;       /* after the last collection                                      */
  uint64_t live_values;
  uint64_t cta_launches, grid_launches, sample_launches;
  double apply_ms;           /* CUDA-event time of the apply kernels of the last run           */
  double sample_ms;          /* ... of build_sampling + sample + histogram compaction          */
  double sample_kernel_ms;   /* ... of the sampling kernel alone                               */
  double norm;               /* squared norm of the final state                                */
  uint64_t distinct_outcomes;
  uint64_t thin_gates;       /* gates applied below the cursor edges without work items (thin_gate) */
  uint64_t h2d_bytes;        /* bytes this handle copied host -> device since jkq_create (every copy is counted) */
  uint64_t d2h_bytes;        /* ... device -> host                                                               */
} jkq_stats;

/* sampling record exported for multi-GPU shot sharding (16 bytes) */
typedef struct {
  uint32_t c0, c1;
  double q0;
} jkq_snode;

/* layout version of the structs in this header; a binding compares it with the value it was written against */
#define JKQ_ABI_VERSION 5u
uint32_t jkq_abi_version(void);

int jkq_create(jkq_handle **out, const jkq_config *cfg);              /* replaces QMDDinit, QMDDpackage.cpp:1248 */
void jkq_destroy(jkq_handle *h);
const char *jkq_last_error(const jkq_handle *h);

int jkq_load_circuit(jkq_handle *h, uint32_t n_qubits, const jkq_op *ops, size_t n_ops);
int jkq_set_option(jkq_handle *h, const char *name, int64_t value);   /* "level_log" = 1: per-gate live counts */

/* apply every op of the loaded circuit to |0..0> (Simulator::AddVariables + ApplyGate per gate) */
int jkq_apply(jkq_handle *h);
/* AssignProbs + MeasureAll x shots (Simulator.cpp:127-224): shots [shot_offset, shot_offset+shots) of
 * the Philox stream `seed`; accumulate = 0 clears the histogram first */
int jkq_sample(jkq_handle *h, uint64_t shots, uint64_t seed, uint64_t shot_offset, int accumulate);
/* jkq_apply + jkq_sample(shots, seed, 0, 0): what `jku_simulator --shots --seed` does */
int jkq_run(jkq_handle *h, uint64_t shots, uint64_t seed);

int jkq_get_counts(jkq_handle *h, uint64_t *keys, uint64_t *counts, size_t cap, size_t *n);
/* give the histogram's device memory back (a 1e9-shot histogram of QFT-32 holds 34 GB); the next jkq_sample allocates anew */
int jkq_release_histogram(jkq_handle *h);
int jkq_get_statevector(jkq_handle *h, uint32_t slot, double *re_im, size_t cap_amplitudes);
int jkq_get_amplitudes(jkq_handle *h, uint64_t first, uint64_t count, double *re_im); /* GetElementOfVector */
/* Amplitudes [first, first+count) of the state a SNAPSHOT op recorded in `slot` (0xFFFFFFFF = the current state).  Up to
 * option "snapshot_dense_max" (26) qubits a snapshot is a dense vector on the device; above, it is a copy of the DIAGRAM
 * (16 B per node) and every amplitude read is one walk through it (GetElementOfVector, Simulator.cpp:415-431) -- the
 * state vector of QASMsimulator.cpp:1012-1030 for a register of any width, read range by range instead of at once. */
int jkq_get_snapshot_amplitudes(jkq_handle *h, uint32_t slot, uint64_t first, uint64_t count, double *re_im);
/* the same with bit b of the index (first + i) standing for qubit qubit_of_bit[b] (a permutation of the n qubits): the
 * reference prints a snapshot's state vector with the first listed qubit as the most significant bit
 * (QASMsimulator.cpp:1016-1028) */
int jkq_get_snapshot_amplitudes_ordered(jkq_handle *h, uint32_t slot, const uint32_t *qubit_of_bit, uint64_t first, uint64_t count,
                                        double *re_im);
int jkq_get_level_counts(jkq_handle *h, uint32_t *per_level, size_t n);              /* Active[v] */
int jkq_get_level_log(jkq_handle *h, uint32_t *out, size_t cap, size_t *n_written);
int jkq_get_stats(jkq_handle *h, jkq_stats *out);
/* P(qubit reads 0) in the current state (Simulator::AssignProbsOne / GetProbability, Simulator.cpp:291-339, 433-464) */
int jkq_get_qubit_probability(jkq_handle *h, uint32_t qubit, double *p0);
/* the 2^k probabilities a PROBABILITIES op recorded in `slot`; *n = 2^k (also when out is NULL or cap too small) */
int jkq_get_probabilities(jkq_handle *h, uint32_t slot, double *out, size_t cap, size_t *n);
/* entries [first, first+count) of the same (streaming read-out of a wide register's probabilities) */
int jkq_get_probabilities_range(jkq_handle *h, uint32_t slot, uint64_t first, uint64_t count, double *out);
/* outcomes of the MEASURE_ONE ops of the last jkq_apply: bit c of *bits = classical bit c, *mask = bits written */
int jkq_get_mid_measurements(jkq_handle *h, uint64_t *bits, uint64_t *mask);

/* multi-GPU shot sharding: export the level-ordered sampling DD, import it on another handle */
int jkq_dd_export(jkq_handle *h, jkq_snode *host_nodes, size_t cap, uint32_t *n_nodes, uint32_t *root);
int jkq_dd_import(jkq_handle *h, uint32_t n_qubits, const jkq_snode *host_nodes, uint32_t n_nodes, uint32_t root,
                  const uint8_t *clbit_of_level);
int jkq_get_measure_map(jkq_handle *h, uint8_t *clbit_of_level, size_t n);
/* the same export / import with DEVICE buffers of the caller (any device of this process): both ends of an NCCL
 * broadcast of the sampling DD, no host staging */
int jkq_dd_export_device(jkq_handle *h, void *dev_nodes, size_t cap, uint32_t *n_nodes, uint32_t *root);
int jkq_dd_import_device(jkq_handle *h, uint32_t n_qubits, const void *dev_nodes, uint32_t n_nodes, uint32_t root,
                         const uint8_t *clbit_of_level);

/* Shot sampling over several GPUs of ONE process (the reference draws its shots one after the other on one
 * core, QASMsimulator.cpp:487-500).  `h` holds the state (jkq_apply or jkq_dd_import); devices[0..n) are CUDA ordinals,
 * h's own device among them or not.  The sampling DD goes to a helper engine per other device by peer copy
 * (cudaMemcpyPeerAsync over NVLink), device i draws shots [i S / n, (i + 1) S / n) of the Philox stream `seed` -- the
 * very shots a one-GPU jkq_sample(h, S, seed, 0, 0) draws -- and the histograms are merged on the device into h's,
 * which jkq_get_counts returns.  An outcome space too large for one GPU stays distributed: st->merged = 0 and
 * st->distinct_total is the sum over the devices. */
typedef struct {
  uint32_t n_devices, merged;
  uint64_t distinct_total;
  uint64_t dd_bytes;              /* bytes of sampling DD copied to every other device                 */
  double broadcast_ms;            /* peer copies, host clock                                           */
  double sample_wall_ms;          /* all devices sampling side by side, host clock                     */
  double sample_kernel_ms_max;    /* CUDA-event time of the sampling kernel, max over the devices      */
  double merge_ms;                /* compaction + peer copies + device merge, host clock               */
  double wall_ms;                 /* the whole call                                                    */
} jkq_multi_stats;
int jkq_sample_multi_gpu(jkq_handle *h, const int32_t *devices, uint32_t n_devices, uint64_t shots, uint64_t seed,
                         jkq_multi_stats *st);

/* Sparse histogram exchange between GPUs: every outcome key has an owner rank (a hash of the key).
 * jkq_counts_partition sorts this handle's (key, count) pairs by owner into DEVICE buffers of the caller (`cap` entries
 * each; NULL = only count) and reports per_owner[world] -- the split sizes of the all-to-all the caller runs (NCCL);
 * jkq_counts_merge adds received pairs (DEVICE buffers) into the handle's histogram (reset = 1: start an empty one
 * with room for `expect` keys), after which jkq_get_counts returns the keys this rank owns.  The caller makes sure
 * the buffers are complete before the call (synchronise the producing stream); both calls return synchronised. */
int jkq_counts_partition(jkq_handle *h, uint32_t world, uint64_t *dev_keys, uint64_t *dev_counts, size_t cap, uint64_t *per_owner);
int jkq_counts_merge(jkq_handle *h, const uint64_t *dev_keys, const uint64_t *dev_counts, size_t n, size_t expect, int reset);

/* Batches of independent circuits (the reference runs one `jku_simulator` process per experiment,
 * sequentially: qasm_simulator_jku.py:340-359).  A pool of `n_engines` small engines; each circuit runs as one
 * persistent single-CTA kernel chain on its own stream, so circuits of the batch occupy different SMs.
 * Circuit i = ops[op_offsets[i] .. op_offsets[i+1]) on n_qubits[i] qubits (GATE / MEASURE / BARRIER ops),
 * `shots` shots of the Philox stream seeds[i]; its outcomes land in keys/counts[i*cap .. i*cap + n_distinct[i]).
 * status[i] != 0: the circuit was not finished on this path (too large for one SM's shared memory, more than
 * `cap` distinct outcomes, ...) and is left to a full jkq_handle.  cfg may be NULL (small pools). */
typedef struct jkq_batch jkq_batch;
int jkq_batch_create(jkq_batch **out, int device, int n_engines, const jkq_config *cfg);
void jkq_batch_destroy(jkq_batch *b);
const char *jkq_batch_last_error(const jkq_batch *b);
int jkq_batch_run(jkq_batch *b, size_t n_circuits, const uint32_t *n_qubits, const jkq_op *ops, const uint64_t *op_offsets,
                  uint64_t shots, const uint64_t *seeds, size_t cap, uint64_t *keys, uint64_t *counts, uint32_t *n_distinct,
                  int32_t *status, uint64_t *gates_applied, double *apply_ms_sum);

/* Per-shot re-simulation (QASMsimulator.cpp:501-522: a circuit with a gate behind a measurement is simulated again
 * for every shot) as ONE kernel: `shots` instances of the same op list run side by side, one CTA per SM, each
 * handling its MEASURE_ONE / RESET / IF ops on the device.  Instance s draws its measurement outcomes from the
 * Philox stream (measure_seed0 + s, op index) -- what jkq_apply draws with option "measure_seed" = measure_seed0 + s --
 * and its final read-out of all measured qubits (MeasureAll) as shot (shot_offset + s) of stream `seed` -- what
 * jkq_sample(1, seed, shot_offset + s) draws.  keys[s] = the outcome, mid_bits[s] (may be NULL) = the classical bits
 * of the MEASURE_ONE ops; status[s] != 0: the instance outgrew the single-CTA path and is left to a full jkq_handle.
 * SNAPSHOT / PROBABILITIES ops record nothing here. */
int jkq_batch_run_shots(jkq_batch *b, uint32_t n_qubits, const jkq_op *ops, size_t n_ops, uint64_t shots, uint64_t seed,
                        uint64_t shot_offset, uint64_t measure_seed0, int physical_reset, uint64_t *keys, uint64_t *mid_bits,
                        int32_t *status, uint64_t *gates_applied, double *kernel_ms);

/* OpenQASM 2.0 front end (host code): text -> the op list jkq_load_circuit takes.  Replaces the executing parser
 * of the reference (src/QASMsimulator.cpp:895-1060, src/QASMscanner.cpp); the semantics it keeps and the points where
 * it deliberately differs are listed at the top of csrc/jkq_qasm.cpp.  `measure` becomes MEASURE_ONE, `reset` RESET,
 * `if` IF, `snapshot(k) q...` a PROBABILITIES op (flag JKQ_QASM_PROBABILITIES, --display_probabilities of main.cpp:84)
 * and, when it lists every qubit, a SNAPSHOT op (flag JKQ_QASM_STATEVECTOR, --display_statevector), both in slot k;
 * `show_probabilities;` a PROBABILITIES op over all qubits in slot JKQ_SLOT_SHOW_PROBABILITIES + i.
 * `include "f";` reads include_dir/f, then ./f; "qelib1.inc" falls back to the built-in standard header.
 * jkq_qasm_parse always sets *out (free it with jkq_program_free); on a syntax error it returns JKQ_E_ARGUMENT and
 * jkq_program_error() gives the message with line and column. */
typedef struct jkq_program jkq_program;
enum { JKQ_QASM_STATEVECTOR = 1, JKQ_QASM_PROBABILITIES = 2 };
#define JKQ_SLOT_SHOW_PROBABILITIES 0x80000000u
typedef struct {
  uint32_t n_qubits, n_clbits;
  size_t n_ops;
  const jkq_op *ops;           /* owned by the program */
  uint32_t n_qregs, n_cregs, n_snapshots;
  int32_t gate_after_measure;  /* a gate follows a measure: the reference re-simulates per shot (QASMsimulator.cpp:501-522) */
  int64_t first_measure;       /* index of the first MEASURE_ONE op, -1 = none */
} jkq_program_info;
int jkq_qasm_parse(const char *text, const char *include_dir, uint32_t flags, jkq_program **out);
const char *jkq_program_error(const jkq_program *p);
void jkq_program_free(jkq_program *p);
int jkq_program_get_info(const jkq_program *p, jkq_program_info *info);
/* register `index` of the qregs (creg = 0) or cregs (creg = 1), in declaration order */
int jkq_program_get_register(const jkq_program *p, int creg, uint32_t index, char *name, size_t name_cap, uint32_t *offset,
                             uint32_t *size);
/* snapshot statement `index` in program order: its slot and its qubits in argument order (first argument = most
 * significant bit of the reference's probability / state-vector index, QASMsimulator.cpp:993-1031) */
int jkq_program_get_snapshot(const jkq_program *p, uint32_t index, uint32_t *slot, uint32_t *qubits, size_t cap, size_t *n_qubits);

/* GPU-measured roofline denominators: uniform random 32-byte sector reads and 64-bit atomicCAS over
 * a `bytes`-sized buffer; returns giga-operations per second */
int jkq_random_sector_bench(jkq_handle *h, uint64_t bytes, int iters, double *gsectors_per_s, double *gcas_per_s);

#ifdef __cplusplus
}
#endif
#endif
