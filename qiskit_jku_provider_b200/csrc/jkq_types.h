// Plain data layouts shared by the CUDA kernels and the host side of the C-ABI library.
//
// Everything the engine keeps lives in HBM behind one `JEng` control block:
//   * node arena      16-byte records, bump-allocated by atomicAdd, two semi-spaces for the
//                     copying collector (replaces QMDDnode/QMDDgetNode, QMDDpackage.h:90-100,
//                     QMDDpackage.cpp:833-860; a state-vector node only needs e[0], e[2] and, after
//                     NormC normalisation, ONE free weight);
//   * unique table    open addressing over 32-bit node ids (replaces Unique[MAXN][NBUCKET],
//                     QMDDpackage.h:235, QMDDutLookup QMDDpackage.cpp:693-763);
//   * value table     tolerance-bucketed reals, 16-byte slots (replaces Ctable/Ctable2/Cmag,
//                     QMDDcomplexD.cpp:45,372-466);
//   * frontier pool   80-byte work items + generation-tagged de-dup tables (replace the recursion
//                     of QMDDmultiply2/QMDDadd and their compute tables, QMDDpackage.cpp:986-1072,
//                     1319-1517).
#ifndef JKQ_TYPES_H
#define JKQ_TYPES_H
#include <stdint.h>

#define JKQ_MAXQ 64

// complex weight id: (R<<32)|I, each half = sign bit + 31-bit index into the value table
// (same encoding as the reference, QMDDcomplex.h:129-131, QMDDcomplexD.cpp:75-83)
#define JW_ZERO 0x0ull
#define JW_ONE 0x0000000100000000ull
#define JW_MONE 0x8000000100000000ull

typedef struct __attribute__((aligned(16))) {
  uint32_t c0;  // child 0 node id | (norm flag << 31).  flag = index of the child whose weight is ONE
  uint32_t c1;  // child 1 node id
  uint64_t w;   // weight id of the OTHER child (may be JW_ZERO -> that child is the terminal 0)
} JNode;

// value-table slot.  (keylo, id1) form one 64-bit word that is claimed by atomicCAS; vbits takes
// atomicMin proposals while the slot is pending.
#define JV_PENDING 0xFFFFFFFFu  // claimed in the current phase, value = min of the proposals so far
#define JV_DROPPED 0xFFFFFFFEu  // claimed once, no entry (the candidate joined a lower neighbour)
#define JV_WAITING 0xFFFFFFFDu  // in-place insertion (real_insert_par): claimed, waits for the owner of the bucket below to finish
#define JV_INF 0x7FF0000000000000ull
typedef struct __attribute__((aligned(16))) {
  unsigned long long vbits;  // non-negative real as raw bits; JV_INF = none
  uint32_t keylo;            // low 32 bits of the bucket number floor(value / tol)
  uint32_t id1;              // 0 = empty, JV_PENDING, JV_DROPPED, JV_WAITING, else (< JV_WAITING) value index + 1
} JVSlot;

// work-item kinds
#define JK_M 1u  // gate x node, at or above the target level
#define JK_F 2u  // fused  P*(wa*A + wb*B) + (I-P)*(wc*C)   below the target, controls remaining
#define JK_S 3u  // plain sum  A + wb*B                      (the reference's QMDDadd pair)
// child-slot kinds
#define JC_ZERO 0u
#define JC_PASS 1u  // finished edge (node id, weight)
#define JC_REF 2u   // result of another work item one level down, times a factor
// item states
#define JS_NEW 0u
#define JS_RETRY 1u
#define JS_DONE 2u
#define JS_DEAD 3u

// flags word of an item: state | chk0 << 2 | chk1 << 4 | hint << 6
typedef struct __attribute__((aligned(16))) {
  uint32_t a, b;     // operand node ids (0 = absent)            -- key, written before publication
  uint32_t tag;      // kind | sel << 4                           -- key
  uint32_t flags;    // owner-written
  uint64_t wa, wb;   // normalised operand weights (kind dependent) -- key
  uint64_t wc;       //                                            -- key
  uint32_t ch[2];    // child-slot target: item index (REF) or node id (PASS)
  uint64_t chw[2];   // child-slot weight: factor (REF) or edge weight (PASS)
  uint64_t rw;       // result edge weight
  uint32_t rnode;    // result edge node
  uint32_t pad;
} JItem;  // 80 bytes = five 16-byte quads

// level roles of a gate (bottom variable = level 0; qubit q lives on level n-1-q,
// reference QASMsimulator.cpp:300)
#define JR_IDLE 0
#define JR_CPOS 1
#define JR_CNEG 2
#define JR_TARGET 3
#define JKQ_PARK_WINDOW 4  // gates looked ahead for JGate.park

typedef struct __attribute__((aligned(16))) {
  double m[8];  // 2x2 matrix, row major, (re, im) pairs -- FP64 entries from the host
  int32_t target;  // level
  int32_t low;     // lowest level carrying a control or the target
  int32_t top;     // highest such level
  int32_t pad;     // index of this gate's matrix among the distinct matrices of the circuit
  int32_t park;    // highest `top` among this gate and the next few: the cursor does not go below it for this gate
  int32_t rsv;
  uint64_t idle;   // bit L: level L carries neither a control nor the target
  uint8_t role[JKQ_MAXQ];
} JGate;  // 160 bytes

// What the cursor warp tells everybody else about the gate it just prepared (or finished).  Two copies,
// indexed by gate parity: the warp may already be writing the next gate's while the others read this one's.
typedef struct {
  int32_t wide_L;    // first level left to everybody (-1: the cursor warp ran the whole gate)
  int32_t wide_low;  // lowest level expanded so far
  uint32_t wide_b;   // first item of level wide_L
  uint32_t lsw;      // first level handled by work items
  uint32_t blocked;  // 1: the cursor sits one level above lsw (jkq_thin.cuh)
  uint32_t collect;  // run the collector after this gate (decided by ONE thread: the pools keep changing)
  uint32_t pad[2];
} JHand;

// sampling record: one per node, level-ordered arena
typedef struct __attribute__((aligned(16))) {
  uint32_t c0, c1;
  double q0;  // probability of taking child 0
} JSNode;

typedef struct {
  // ---- configuration (host-written before launch) ----
  uint32_t n;
  uint32_t node_cap, utab_mask, val_cap, vtab_mask, pend_cap, item_cap, dtab_mask;
  uint32_t cta_item_limit;  // CTA mode bails out to grid mode above this frontier size
  uint32_t gc_min;          // never collect below this many allocated nodes
  uint32_t cta_node_limit;  // testing aid: cap on the shared-memory node pool of the CTA path (0 = its real size)
  uint32_t cta_val_limit;   // testing aid: the same for the shared-memory value table
  uint32_t cta_wide_ok;     // a one-warp single-CTA launch may ask for the wide one (want_wide) when a level outgrows the warp
  uint32_t thin_off;        // testing aid: bit 0 no thin gates with controls above the target, 1 none with controls below, 2 none wider than one lane, 3 none at all,
                            // 4 no parked cursor, 5 no thin_probe, 6 no chain sums (thin_probe_sums)
  uint32_t solo_max;       // widest level the cursor warp runs by itself (one lane per item); 0 = JKQ_WARP.  Wider levels are shared out
                           // over the warps of the CTA / grid (run_level)
  uint32_t sums_ask;       // the single-CTA kernel without thin_probe_sums counts the gates that would have used it down from here;
                           // at 1 it stops and asks for the kernel that has it (want_sums).  0: such gates take the work items
  uint32_t cta_no_reserve; // testing aid: no reserve in the shared pools, so the root rebuild of an abandoned gate fails too (saved-cursor hand-over)
  double tol;
  double inv_tol;
  // ---- arrays ----
  JNode *nodes, *nodes_alt;
  unsigned long long *utab;  // generation << 32 | node id: a collection starts a new generation instead of clearing
  double *vals, *vals_alt;
  JVSlot *vtab;
  uint32_t *vbm;      // one bit per 128-byte line (8 slots) of the HBM value table: some slot of the line was claimed since
                      // the last sweep -- a sweep clears those lines instead of the whole table (4.3 GB at 2^26 reals)
  uint32_t *pend;     // value-table slots claimed in the current phase
  JItem *items;
  unsigned long long *dtab;  // F/S item de-dup: generation << 32 | item index + 1
  unsigned long long *mdir;  // M item de-dup, indexed by node id: generation << 32 | item index + 1
  uint32_t *fwd;      // node_cap entries: forwarding ids for the collector
  uint32_t *vfwd;     // val_cap entries: value id remap for the collector
  uint32_t *lvl_log;  // optional [gate][level] live-node counts
  void *thin_spill;   // the cursor (ThinState) of a single-CTA kernel that could not rebuild up to the root
  uint64_t *mat_u;      // [n distinct matrices][4] cached canonical ids of gate-matrix entries
  uint32_t *mat_epoch;  // [n distinct matrices] val_epoch at which mat_u was filled (0 = never)
  // ---- mutable state ----
  uint32_t node_count;  // next free node id (id 0 = terminal)
  uint32_t val_count;
  uint32_t item_count;
  uint32_t pend_n[3];  // one per barrier-ring slot
  uint32_t gen;  // de-dup generation (one per gate attempt)
  uint32_t root;
  uint64_t rootw;
  uint64_t cur_u[4];  // canonical ids of the current gate's matrix (device-written at gate setup)
  JHand hand[2];  // per-gate hand-over from the cursor warp to everybody else, by gate parity
  uint32_t lb[JKQ_MAXQ + 1], le[JKQ_MAXQ + 1];  // item ranges of the levels of the current gate (grid path)
  uint32_t gc_limit;
  uint32_t gc_count;   // collector cursor
  uint32_t val_gc_limit;  // sweep the reals only above this many (reference complex_limit)
  uint32_t val_epoch;     // bumped whenever a sweep renumbers the reals
  uint32_t ut_gen;        // generation of the unique-table entries that count (older ones read as empty)
  uint32_t vt_cur;        // mask of the part of the value table in use (= vtab_mask: see vt_mask_for, jkq_apply.cuh)
  uint32_t vt_dirty_all;  // the value table has never been cleared by this engine: the next sweep clears all of it
  double norm;        // squared norm of the state at the last finalisation
  uint32_t lvl_off[JKQ_MAXQ + 2];  // after a collection: nodes of level L are [lvl_off[L], lvl_off[L] + lvl_cnt[L])
  uint32_t lvl_cnt[JKQ_MAXQ];      // live nodes per level at the last collection
  // ---- status ----
  int32_t status;       // 0 ok, >0 error code
  int32_t next_gate;    // first gate not yet applied
  int32_t want_grid;    // CTA mode asked for the grid path
  uint32_t tables_valid;  // HBM unique / value hash tables match the arena (MODE_CTA lets them go stale)
  uint32_t cursor_saved;  // thin_spill holds the state; root / rootw are stale; only the grid path may continue
  uint32_t want_sums;     // raised together with want_grid: relaunch the single-CTA kernel that has thin_probe_sums
  uint32_t want_wide;     // raised together with want_grid by a one-warp launch: the wide single-CTA launch is enough
  // ---- counters ----
  unsigned long long st[12];  // JST_* below
  // ---- grid barrier ----
  unsigned int bar_count;
  unsigned int bar_gen;
  unsigned int bar_flags[3];
  unsigned long long prof[16];  // JKQ_PROFILE builds: cycle counters, see jkq_b200.cu jkq_get_stats
  unsigned long long prof_x[8];   // JKQ_PROFILE, grid path: cycles / count of item loops: [0,1] first run, no miss  [2,3] first run with a value miss  [4,5] reruns
  unsigned long long prof_w[32];  // JKQ_PROFILE, grid path: level phases by width class (log4 of the item count): [c] cycles, [8+c] levels, [16+c] barriers, [24+c] items
  uint32_t gc_bsum[1024];  // collector: live reals per CTA chunk
  unsigned int dbg_timeout_block;   // debugging aid: who timed out at a grid barrier, and where everyone was
  unsigned int dbg_phase[1024];
} JEng;

// counters (JEng.st; MODE_CTA keeps 32-bit copies in shared memory and adds them in when it leaves)
#define JST_ITEMS 0      // frontier pairs processed (the reference's CTlook, QMDDpackage.cpp:993)
#define JST_NODES 1      // unique-table inserts
#define JST_VALS 2       // reals appended to the value table
#define JST_RETRY 3      // value-commit rounds
#define JST_THIN 4       // gates applied by thin_gate (no work items)
#define JST_GATES 5
#define JST_MAXF 6       // largest per-gate item count (max, not sum)
#define JST_GC_RUNS 7
#define JST_GC_COPIED 8
#define JST_N 9

// error codes (JEng.status, also returned by the C-ABI)
#define JKQ_OK 0
#define JKQ_ERR_NODE_FULL 1
#define JKQ_ERR_UTAB_FULL 2
#define JKQ_ERR_VAL_FULL 3
#define JKQ_ERR_ITEM_FULL 4
#define JKQ_ERR_DTAB_FULL 5
#define JKQ_ERR_PEND_FULL 6
#define JKQ_ERR_BARRIER_TIMEOUT 7
#define JKQ_ERR_ZERO_STATE 8
#define JKQ_ERR_INTERNAL 9
#define JKQ_ERR_RETRY_LIMIT 10  // a level kept meeting new values for JKQ_RETRY_LIMIT rounds (watchdog)
#define JKQ_ERR_VAL_RANGE 11  // a weight of absurd magnitude (>= 1e8) or NaN: numerical breakdown, not a capacity problem
#define JKQ_RETRY_LIMIT 512u

#endif
