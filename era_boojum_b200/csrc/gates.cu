// Per-row gate / quotient evaluator over general-purpose columns and over specialised columns.
//
// Reference semantics (what is computed, bit-exact):
//   * a gate's constraint terms come from GateConstraintEvaluator::evaluate_once (src/cs/traits/evaluator.rs:145-152),
//     repeated `num_repetitions` times per row with PerChunkOffset (RowwiseEvaluator, evaluator.rs:376-397); row-shared
//     constants are read once at repetition 0 (load_row_shared_constants);
//   * every pushed term is folded with the next alpha power, the power index running on across the gates of the row
//     (push_evaluation_result, src/cs/implementations/buffering_source.rs:304-362), the per-gate sum is multiplied by the
//     gate's selector and added to the quotient (proceed_to_next_gate, :158-221);
//   * the selector of a gate is the product along its path in the selector tree of const_i or (1 - const_i)
//     (compute_selector_subpath, src/cs/implementations/prover.rs:2775-2916); the gate's own constants start at column
//     `path length` (constant_placement_offset, prover.rs:1000-1013);
//   * driver: the row loop of prove_cpu_basic (prover.rs:1031-1080) over the first Q cosets of the LDE.
//   * gates placed on SPECIALISED columns (GatePlacementStrategy::UseSpecializedColumns, prover.rs:653-801) own a fixed range
//     of columns: no selector, the first repetition starts at the gate's initial offset, constants start behind those of
//     the general-purpose gates (shared by the repetitions when share_constants); their terms precede the general-purpose ones.
// The gate itself is data: the SSA program the reference's own GPU hook records (gpu_synthesizer::GPUDataCapture,
// src/gpu_synthesizer/mod.rs:115-133, 354-443) - Index::{VariablePoly, WitnessPoly, ConstantPoly, TemporaryValue,
// ConstantValue} and Relation::{Add, Double, Sub, Negate, Mul, Square, Inverse} - so any evaluator the reference can
// capture runs here unchanged.  One thread owns one (coset, row) point and interprets the programs; the row's columns
// are read once (coalesced across threads), temporaries live in thread-local memory.
#include <vector>
#include "ctx.hpp"

namespace bj {

constexpr int GATE_MAX_TMP = 128;            // live temporaries per thread after host-side slot allocation
constexpr u32 GATE_MAX_PROGRAM_TMP = 1u << 20;  // temporaries a recorded program may name (SSA: one per relation)
constexpr u32 GATE_OP_PUSH = 7;               // internal: fold operand a into the accumulator with alpha power `dst` of the repetition

struct DevOperand {
  u32 kind;  // bj_gate_index kinds
  u32 pad;
  u64 value;
};
struct DevOp {
  u32 op;
  u32 dst;
  DevOperand a, b;
};
struct DevGate {
  u32 ops_begin, n_ops;
  u32 n_writes;                               // quotient terms per repetition
  u32 num_repetitions;
  u32 var_offset, wit_offset, const_offset;  // PerChunkOffset
  u32 var_base, wit_base;                     // first column of repetition 0 (specialised placement; 0 for general purpose)
  u32 const_placement;                        // first constant column of the gate (= selector path length)
  u32 path_len;
  u32 path_bits;  // bit i = path[i]
  u32 term_base;  // index of the gate's first alpha power
};

struct GateEvalParams {
  const DevGate* gates;
  u32 n_gates;
  const DevOp* ops;
  const u64* const* vars;
  const u64* const* wits;
  const u64* const* consts;
  const u64* alphas;  // (c0, c1) per term
  u64 n_rows;
  u64* q_c0;
  u64* q_c1;
};

__device__ __forceinline__ u64 gate_fetch(const DevOperand& o, const u64* tmp, const GateEvalParams& p, u64 t, u32 vbase,
                                          u32 wbase, u32 cbase, u32 cshared) {
  switch (o.kind) {
    case BJ_IDX_VARIABLE: return p.vars[vbase + (u32)o.value][t];
    case BJ_IDX_WITNESS: return p.wits[wbase + (u32)o.value][t];
    case BJ_IDX_CONSTANT_POLY: return p.consts[cbase + (u32)o.value][t];
    case BJ_IDX_CONSTANT_POLY_SHARED: return p.consts[cshared + (u32)o.value][t];
    case BJ_IDX_TEMPORARY: return tmp[(u32)o.value];
    default: return o.value;  // BJ_IDX_CONSTANT_VALUE
  }
}

__global__ void __launch_bounds__(128) gate_eval_kernel(const GateEvalParams p) {
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p.n_rows) return;
  u64 tmp[GATE_MAX_TMP];
  gl::e2 q = {0, 0};
  for (u32 g = 0; g < p.n_gates; g++) {
    const DevGate gate = p.gates[g];
    gl::e2 acc = {0, 0};
    for (u32 rep = 0; rep < gate.num_repetitions; rep++) {
      const u32 vbase = gate.var_base + rep * gate.var_offset, wbase = gate.wit_base + rep * gate.wit_offset;
      const u32 cshared = gate.const_placement, cbase = cshared + rep * gate.const_offset;
      const u64* alpha_rep = p.alphas + 2 * (size_t)(gate.term_base + rep * gate.n_writes);
      for (u32 i = 0; i < gate.n_ops; i++) {
        const DevOp op = p.ops[gate.ops_begin + i];
        const u64 a = gate_fetch(op.a, tmp, p, t, vbase, wbase, cbase, cshared);
        u64 r;
        switch (op.op) {
          case BJ_REL_ADD: r = gl::add_lazy(a, gate_fetch(op.b, tmp, p, t, vbase, wbase, cbase, cshared)); break;
          case BJ_REL_DOUBLE: r = gl::add_lazy(a, a); break;
          case BJ_REL_SUB: r = gl::sub_lazy(a, gate_fetch(op.b, tmp, p, t, vbase, wbase, cbase, cshared)); break;
          case BJ_REL_NEGATE: r = gl::neg(a); break;
          case BJ_REL_MUL: r = gl::mul(a, gate_fetch(op.b, tmp, p, t, vbase, wbase, cbase, cshared)); break;
          case BJ_REL_SQUARE: r = gl::sqr(a); break;
          case BJ_REL_INVERSE: r = gl_inv_chain(gl::canon(a)); break;
          default: {  // GATE_OP_PUSH: push_evaluation_result - the term times its alpha power goes into the gate's accumulator
            const u64 a0 = __ldg(alpha_rep + 2 * op.dst), a1 = __ldg(alpha_rep + 2 * op.dst + 1);
            acc.c0 = gl::add(acc.c0, gl::mul(a, a0));
            acc.c1 = gl::add(acc.c1, gl::mul(a, a1));
            continue;
          }
        }
        tmp[op.dst] = r;
      }
    }
    u64 sel = 1;
    for (u32 i = 0; i < gate.path_len; i++) {
      const u64 c = gl::canon(p.consts[i][t]);
      sel = gl::mul(sel, ((gate.path_bits >> i) & 1) ? c : gl::canon(gl::sub(1, c)));
    }
    q.c0 = gl::add(q.c0, gl::mul(acc.c0, sel));
    q.c1 = gl::add(q.c1, gl::mul(acc.c1, sel));
  }
  p.q_c0[t] = gl::canon(gl::add(p.q_c0[t], gl::canon(q.c0)));
  p.q_c1[t] = gl::canon(gl::add(p.q_c1[t], gl::canon(q.c1)));
}

}  // namespace bj

using namespace bj;

extern "C" int32_t bj_quotient_gates_general_purpose(bj_ctx* ctx, const bj_gate_desc* h_gates, uint32_t n_gates,
                                                     const uint64_t* const* h_variable_cols, uint32_t n_variables,
                                                     const uint64_t* const* h_witness_cols, uint32_t n_witnesses,
                                                     const uint64_t* const* h_constant_cols, uint32_t n_constants,
                                                     const uint64_t* h_alpha_powers, uint32_t n_alpha_powers,
                                                     uint64_t n_points, uint64_t* d_q_c0, uint64_t* d_q_c1) {
  bj::DeviceGuard device_guard(ctx);
  if (!ctx || !h_gates || n_gates == 0 || !d_q_c0 || !d_q_c1 || n_points == 0 || (!h_alpha_powers && n_alpha_powers))
    BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "bj_quotient_gates_general_purpose: bad argument");
  std::vector<DevGate> gates;
  std::vector<DevOp> ops;
  uint64_t total_terms = 0;
  // operand range checks against the columns the caller passed (all repetitions)
  auto check_index = [&](const bj_gate_index& ix, const bj_gate_desc& g, DevOperand* out) -> bool {
    const uint32_t reps = g.num_repetitions ? g.num_repetitions - 1 : 0;
    switch (ix.kind) {
      case BJ_IDX_VARIABLE:
        if (g.variables_initial_offset + ix.value + (uint64_t)reps * g.variables_offset >= n_variables) return false;
        break;
      case BJ_IDX_WITNESS:
        if (g.witnesses_initial_offset + ix.value + (uint64_t)reps * g.witnesses_offset >= n_witnesses) return false;
        break;
      case BJ_IDX_CONSTANT_POLY:
        if (g.constants_placement_offset + ix.value + (uint64_t)reps * g.constants_offset >= n_constants) return false;
        break;
      case BJ_IDX_CONSTANT_POLY_SHARED:
        if (g.constants_placement_offset + ix.value >= n_constants) return false;
        break;
      case BJ_IDX_TEMPORARY:
        if (ix.value >= GATE_MAX_PROGRAM_TMP) return false;
        break;
      case BJ_IDX_CONSTANT_VALUE: break;
      default: return false;
    }
    out->kind = ix.kind;
    out->pad = 0;
    out->value = ix.kind == BJ_IDX_CONSTANT_VALUE ? gl::canon(ix.value) : ix.value;
    return true;
  };
  for (uint32_t gi = 0; gi < n_gates; gi++) {
    const bj_gate_desc& g = h_gates[gi];
    if (g.selector_path_len > 32 || g.selector_path_len > n_constants || (g.n_relations && !g.relations) ||
        (g.n_writes && !g.writes))
      BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "gate descriptor: bad selector path or NULL program");
    DevGate d{};
    d.ops_begin = (u32)ops.size();
    d.n_writes = g.n_writes;
    d.num_repetitions = g.num_repetitions;
    d.var_offset = g.variables_offset;
    d.wit_offset = g.witnesses_offset;
    d.var_base = g.variables_initial_offset;
    d.wit_base = g.witnesses_initial_offset;
    d.const_offset = g.constants_offset;
    d.const_placement = g.constants_placement_offset;
    d.path_len = g.selector_path_len;
    d.path_bits = 0;
    d.term_base = (u32)total_terms;
    for (uint32_t i = 0; i < g.selector_path_len; i++)
      if (g.selector_path[i]) d.path_bits |= 1u << i;
    // 1. the recorded program (SSA as GPUVariablesContext records it: every relation defines a fresh TemporaryValue, used
    //    only afterwards), with the pushes of the quotient terms placed right behind the relation that defines them
    //    (push_evaluation_result is called inline by evaluate_once; the alpha power of a term is fixed by its write index)
    std::vector<DevOp> prog;
    prog.reserve(g.n_relations + g.n_writes);
    std::vector<int32_t> def_at;  // program temporary -> index in `prog` of its defining op (-1: undefined)
    auto defined = [&](const bj_gate_index& ix) { return ix.kind != BJ_IDX_TEMPORARY || (ix.value < def_at.size() && def_at[ix.value] >= 0); };
    std::vector<std::vector<uint32_t>> pushes_of(g.n_relations);  // relation index -> write indices it feeds
    std::vector<uint32_t> late_pushes;                              // writes of bare columns / constants
    for (uint32_t i = 0; i < g.n_relations; i++) {
      const bj_gate_relation& r = g.relations[i];
      if (r.op > BJ_REL_INVERSE || r.dst_temporary >= GATE_MAX_PROGRAM_TMP)
        BJ_FAIL(ctx, BJ_ERR_UNSUPPORTED, "gate program: unknown relation or temporary index beyond 2^20");
      if (r.dst_temporary >= def_at.size()) def_at.resize(r.dst_temporary + 1, -1);
    }
    {
      std::vector<int32_t> def_rel(def_at.size(), -1);
      for (uint32_t i = 0; i < g.n_relations; i++) {
        if (def_rel[g.relations[i].dst_temporary] >= 0)
          BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "gate program: temporary defined twice (programs are SSA)");
        def_rel[g.relations[i].dst_temporary] = (int32_t)i;
      }
      for (uint32_t k = 0; k < g.n_writes; k++) {
        const bj_gate_index& w = g.writes[k];
        if (w.kind == BJ_IDX_TEMPORARY) {
          if (w.value >= def_rel.size() || def_rel[w.value] < 0) BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "gate program: write of an undefined temporary");
          pushes_of[def_rel[w.value]].push_back(k);
        } else {
          late_pushes.push_back(k);
        }
      }
    }
    auto emit_push = [&](uint32_t k) -> bool {
      DevOp o{};
      o.op = GATE_OP_PUSH;
      o.dst = k;
      return check_index(g.writes[k], g, &o.a) && (prog.push_back(o), true);
    };
    for (uint32_t i = 0; i < g.n_relations; i++) {
      const bj_gate_relation& r = g.relations[i];
      DevOp o{};
      o.op = r.op;
      o.dst = r.dst_temporary;
      const bool binary = r.op == BJ_REL_ADD || r.op == BJ_REL_SUB || r.op == BJ_REL_MUL;
      const bj_gate_index bdummy{BJ_IDX_CONSTANT_VALUE, 0, 0};
      if (!defined(r.a) || (binary && !defined(r.b)) || !check_index(r.a, g, &o.a) || !check_index(binary ? r.b : bdummy, g, &o.b))
        BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "gate program: operand out of range or temporary used before definition");
      def_at[r.dst_temporary] = (int32_t)prog.size();
      prog.push_back(o);
      for (uint32_t k : pushes_of[i])
        if (!emit_push(k)) BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "gate program: write operand out of range");
    }
    for (uint32_t k : late_pushes)
      if (!emit_push(k)) BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "gate program: write operand out of range");
    // 2. slot allocation: a temporary lives from its definition to its last use; its slot is then reused (the kernel reads both
    //    operands before it writes the destination, so a destination may take over the slot of an operand that dies there)
    {
      std::vector<int32_t> last_use(def_at.size(), -1), slot(def_at.size(), -1);
      for (size_t i = 0; i < prog.size(); i++) {
        if (prog[i].a.kind == BJ_IDX_TEMPORARY) last_use[prog[i].a.value] = (int32_t)i;
        if (prog[i].op != GATE_OP_PUSH && prog[i].b.kind == BJ_IDX_TEMPORARY) last_use[prog[i].b.value] = (int32_t)i;
      }
      std::vector<uint32_t> free_slots;
      uint32_t next_slot = 0;
      for (size_t i = 0; i < prog.size(); i++) {
        DevOp& o = prog[i];
        const bool is_push = o.op == GATE_OP_PUSH;
        uint64_t ta = o.a.kind == BJ_IDX_TEMPORARY ? o.a.value : ~0ull, tb = (!is_push && o.b.kind == BJ_IDX_TEMPORARY) ? o.b.value : ~0ull;
        if (ta != ~0ull) o.a.value = (uint64_t)slot[ta];
        if (tb != ~0ull) o.b.value = (uint64_t)slot[tb];
        if (ta != ~0ull && last_use[ta] == (int32_t)i) free_slots.push_back((uint32_t)slot[ta]);
        if (tb != ~0ull && tb != ta && last_use[tb] == (int32_t)i) free_slots.push_back((uint32_t)slot[tb]);
        if (is_push) continue;
        const uint32_t t_dst = o.dst;
        uint32_t sl;
        if (!free_slots.empty()) {
          sl = free_slots.back();
          free_slots.pop_back();
        } else {
          sl = next_slot++;
        }
        if (sl >= (uint32_t)GATE_MAX_TMP)
          BJ_FAIL(ctx, BJ_ERR_UNSUPPORTED, "gate program: more than 128 temporaries live at once");
        slot[t_dst] = (int32_t)sl;
        o.dst = sl;
        if (last_use[t_dst] < 0) free_slots.push_back(sl);  // defined but never read
      }
    }
    d.n_ops = (u32)prog.size();
    ops.insert(ops.end(), prog.begin(), prog.end());
    total_terms += (uint64_t)g.n_writes * g.num_repetitions;
    gates.push_back(d);
  }
  if (total_terms > n_alpha_powers) BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "not enough alpha powers for the gate terms");
  std::vector<u64> alphas(2 * (size_t)total_terms);
  for (size_t i = 0; i < alphas.size(); i++) alphas[i] = gl::canon(h_alpha_powers[i]);
  GateEvalParams p{};
  void* d;
  BJ_TRY(param_upload(ctx, gates.data(), sizeof(DevGate) * gates.size(), &d));
  p.gates = (const DevGate*)d;
  p.n_gates = n_gates;
  static const DevOp dummy_op{};
  // small programs ride in the parameter arena; a long one (the Poseidon2 flattened gate is ~9k relations) gets its own
  // stream-ordered buffer, released behind the kernel
  void* big_program = nullptr;
  const size_t ops_bytes = sizeof(DevOp) * std::max<size_t>(ops.size(), 1);
  if (ops_bytes > (128u << 10)) {
    BJ_CUDA(ctx, cudaMallocAsync(&big_program, ops_bytes, ctx->stream));
    const cudaError_t e = cudaMemcpyAsync(big_program, ops.data(), ops_bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) {
      cudaFreeAsync(big_program, ctx->stream);
      BJ_FAIL(ctx, BJ_ERR_CUDA, std::string("gate program upload: ") + cudaGetErrorString(e));
    }
    d = big_program;
  } else {
    BJ_TRY(param_upload(ctx, ops.empty() ? &dummy_op : ops.data(), ops_bytes, &d));
  }
  struct ProgramGuard {  // freed (stream-ordered, i.e. after the kernel) on every exit path
    void* p;
    cudaStream_t s;
    ~ProgramGuard() {
      if (p) cudaFreeAsync(p, s);
    }
  } program_guard{big_program, ctx->stream};
  p.ops = (const DevOp*)d;
  static const u64* const null_ptr = nullptr;
  BJ_TRY(param_upload(ctx, n_variables ? (const void*)h_variable_cols : (const void*)&null_ptr, sizeof(u64*) * std::max(n_variables, 1u), &d));
  p.vars = (const u64* const*)d;
  BJ_TRY(param_upload(ctx, n_witnesses ? (const void*)h_witness_cols : (const void*)&null_ptr, sizeof(u64*) * std::max(n_witnesses, 1u), &d));
  p.wits = (const u64* const*)d;
  BJ_TRY(param_upload(ctx, n_constants ? (const void*)h_constant_cols : (const void*)&null_ptr, sizeof(u64*) * std::max(n_constants, 1u), &d));
  p.consts = (const u64* const*)d;
  static const u64 zero2[2] = {0, 0};
  BJ_TRY(param_upload(ctx, alphas.empty() ? (const void*)zero2 : (const void*)alphas.data(), sizeof(u64) * std::max<size_t>(alphas.size(), 2), &d));
  p.alphas = (const u64*)d;
  p.n_rows = n_points;
  p.q_c0 = (u64*)d_q_c0;
  p.q_c1 = (u64*)d_q_c1;
  gate_eval_kernel<<<(unsigned)((n_points + 127) / 128), 128, 0, ctx->stream>>>(p);
  BJ_LAUNCH_CHECK(ctx);
  return BJ_OK;
}
