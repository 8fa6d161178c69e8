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

constexpr int GATE_MAX_TMP = 96;

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
  u32 writes_begin, n_writes;
  u32 num_repetitions;
  u32 var_offset, wit_offset, const_offset;  // PerChunkOffset
  u32 var_base, wit_base;                     // first column of repetition 0 (specialised placement; 0 for general purpose)
  u32 const_placement;                        // first constant column of the gate (= selector path length)
  u32 path_len;
  u32 path_bits;  // bit i = path[i]
};

struct GateEvalParams {
  const DevGate* gates;
  u32 n_gates;
  const DevOp* ops;
  const DevOperand* writes;
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
  u32 term = 0;
  for (u32 g = 0; g < p.n_gates; g++) {
    const DevGate gate = p.gates[g];
    gl::e2 acc = {0, 0};
    for (u32 rep = 0; rep < gate.num_repetitions; rep++) {
      const u32 vbase = gate.var_base + rep * gate.var_offset, wbase = gate.wit_base + rep * gate.wit_offset;
      const u32 cshared = gate.const_placement, cbase = cshared + rep * gate.const_offset;
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
          default: r = gl_inv_chain(gl::canon(a)); break;  // BJ_REL_INVERSE
        }
        tmp[op.dst] = r;
      }
      for (u32 i = 0; i < gate.n_writes; i++) {
        const u64 v = gate_fetch(p.writes[gate.writes_begin + i], tmp, p, t, vbase, wbase, cbase, cshared);
        const u64 a0 = __ldg(p.alphas + 2 * term), a1 = __ldg(p.alphas + 2 * term + 1);
        acc.c0 = gl::add(acc.c0, gl::mul(v, a0));
        acc.c1 = gl::add(acc.c1, gl::mul(v, a1));
        term++;
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
  std::vector<DevOperand> writes;
  uint64_t total_terms = 0;
  auto check_index = [&](const bj_gate_index& ix, const bj_gate_desc& g, uint32_t n_tmp_defined, DevOperand* out) -> bool {
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
        if (ix.value >= GATE_MAX_TMP || ix.value >= n_tmp_defined) return false;
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
    d.n_ops = g.n_relations;
    d.writes_begin = (u32)writes.size();
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
    for (uint32_t i = 0; i < g.selector_path_len; i++)
      if (g.selector_path[i]) d.path_bits |= 1u << i;
    // temporaries must be defined before use (SSA order, as GPUVariablesContext records them)
    bool defined[GATE_MAX_TMP] = {false};
    auto tmp_ok = [&](const bj_gate_index& ix) { return ix.kind != BJ_IDX_TEMPORARY || (ix.value < GATE_MAX_TMP && defined[ix.value]); };
    uint32_t max_tmp = 0;
    for (uint32_t i = 0; i < g.n_relations; i++) {
      const bj_gate_relation& r = g.relations[i];
      if (r.op > BJ_REL_INVERSE || r.dst_temporary >= GATE_MAX_TMP)
        BJ_FAIL(ctx, BJ_ERR_UNSUPPORTED, "gate program: unknown relation or more than 96 temporaries");
      DevOp o{};
      o.op = r.op;
      o.dst = r.dst_temporary;
      const bool binary = r.op == BJ_REL_ADD || r.op == BJ_REL_SUB || r.op == BJ_REL_MUL;
      bj_gate_index bdummy{BJ_IDX_CONSTANT_VALUE, 0, 0};
      if (!tmp_ok(r.a) || (binary && !tmp_ok(r.b)) || !check_index(r.a, g, GATE_MAX_TMP, &o.a) ||
          !check_index(binary ? r.b : bdummy, g, GATE_MAX_TMP, &o.b))
        BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "gate program: operand out of range or temporary used before definition");
      defined[r.dst_temporary] = true;
      if (r.dst_temporary + 1 > max_tmp) max_tmp = r.dst_temporary + 1;
      ops.push_back(o);
    }
    for (uint32_t i = 0; i < g.n_writes; i++) {
      DevOperand w;
      if (!tmp_ok(g.writes[i]) || !check_index(g.writes[i], g, max_tmp ? max_tmp : 1, &w))
        BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "gate program: write operand out of range");
      writes.push_back(w);
    }
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
  BJ_TRY(param_upload(ctx, ops.empty() ? &dummy_op : ops.data(), sizeof(DevOp) * std::max<size_t>(ops.size(), 1), &d));
  p.ops = (const DevOp*)d;
  static const DevOperand dummy_w{};
  BJ_TRY(param_upload(ctx, writes.empty() ? &dummy_w : writes.data(), sizeof(DevOperand) * std::max<size_t>(writes.size(), 1), &d));
  p.writes = (const DevOperand*)d;
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
