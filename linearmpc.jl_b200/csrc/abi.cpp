// lmpc_gpu_abi_layout: sizeof / offsetof of every struct that crosses the C ABI, as compiled
// into this library.  Bindings in other languages are checked against it (tests/test_abi_cpu.py
// parses the Julia structs; LinearMPCGPU.jl checks itself when the module loads).
#include <cstddef>
#include <cstdint>
#include <vector>
#include "lmpc_gpu.h"

namespace {
struct Rec { std::vector<int32_t> v; };
#define F(T, f) (int32_t)offsetof(T, f)
void emit(std::vector<int32_t> &o, int id, size_t size, std::initializer_list<int32_t> offs) {
  o.push_back(id); o.push_back((int32_t)size); o.push_back((int32_t)offs.size());
  for (int32_t x : offs) o.push_back(x);
}
}  // namespace

extern "C" int lmpc_gpu_abi_layout(int32_t *out, int32_t cap) {
  std::vector<int32_t> o;
  emit(o, LMPC_ABI_PROBLEM, sizeof(lmpc_problem),
       {F(lmpc_problem, n), F(lmpc_problem, m), F(lmpc_problem, ms), F(lmpc_problem, nth), F(lmpc_problem, nu),
        F(lmpc_problem, nx), F(lmpc_problem, H), F(lmpc_problem, f), F(lmpc_problem, f_theta), F(lmpc_problem, A),
        F(lmpc_problem, bu), F(lmpc_problem, bl), F(lmpc_problem, W), F(lmpc_problem, senses), F(lmpc_problem, K),
        F(lmpc_problem, soft_weight), F(lmpc_problem, has_binaries), F(lmpc_problem, is_symmetric),
        F(lmpc_problem, n_break_points)});
  emit(o, LMPC_ABI_SETTINGS, sizeof(lmpc_settings),
       {F(lmpc_settings, primal_tol), F(lmpc_settings, dual_tol), F(lmpc_settings, zero_tol),
        F(lmpc_settings, progress_tol), F(lmpc_settings, fval_bound), F(lmpc_settings, cycle_tol),
        F(lmpc_settings, iter_limit), F(lmpc_settings, kernel), F(lmpc_settings, certified_iterations)});
  emit(o, LMPC_ABI_SIM_EXT, sizeof(lmpc_sim_ext),
       {F(lmpc_sim_ext, Rtraj), F(lmpc_sim_ext, Dtraj), F(lmpc_sim_ext, Ft), F(lmpc_sim_ext, Gt),
        F(lmpc_sim_ext, Gdt), F(lmpc_sim_ext, C), F(lmpc_sim_ext, Dd), F(lmpc_sim_ext, K), F(lmpc_sim_ext, Vtraj),
        F(lmpc_sim_ext, ny), F(lmpc_sim_ext, Xhat_end), F(lmpc_sim_ext, rstride), F(lmpc_sim_ext, dstride)});
  emit(o, LMPC_ABI_CONSTRAINT, sizeof(lmpc_constraint),
       {F(lmpc_constraint, rows), F(lmpc_constraint, Au), F(lmpc_constraint, Ax), F(lmpc_constraint, Ar),
        F(lmpc_constraint, Ad), F(lmpc_constraint, Aup), F(lmpc_constraint, ub), F(lmpc_constraint, lb),
        F(lmpc_constraint, ks), F(lmpc_constraint, nks), F(lmpc_constraint, soft), F(lmpc_constraint, prio),
        F(lmpc_constraint, Ap)});
  emit(o, LMPC_ABI_MODEL, sizeof(lmpc_model),
       {F(lmpc_model, nx), F(lmpc_model, nu), F(lmpc_model, ny), F(lmpc_model, nd), F(lmpc_model, Np),
        F(lmpc_model, Nc), F(lmpc_model, F), F(lmpc_model, G), F(lmpc_model, C), F(lmpc_model, Gd),
        F(lmpc_model, Dd), F(lmpc_model, Q), F(lmpc_model, R), F(lmpc_model, Rr), F(lmpc_model, Qf),
        F(lmpc_model, umin), F(lmpc_model, umax), F(lmpc_model, constraints), F(lmpc_model, n_constraints),
        F(lmpc_model, reference_tracking), F(lmpc_model, preprocess), F(lmpc_model, move_blocks),
        F(lmpc_model, n_move_blocks), F(lmpc_model, reference_preview), F(lmpc_model, disturbance_preview),
        F(lmpc_model, K), F(lmpc_model, np), F(lmpc_model, Ex), F(lmpc_model, ex), F(lmpc_model, Eu),
        F(lmpc_model, eu), F(lmpc_model, parameter_preview),
        F(lmpc_model, f_offset), F(lmpc_model, h_offset), F(lmpc_model, xo), F(lmpc_model, uo),
        F(lmpc_model, reference_condensation)});
  emit(o, LMPC_ABI_SIM_STEPPED, sizeof(lmpc_sim_stepped),
       {F(lmpc_sim_stepped, dynamics), F(lmpc_sim_stepped, user), F(lmpc_sim_stepped, Ft), F(lmpc_sim_stepped, Gt),
        F(lmpc_sim_stepped, Gdt), F(lmpc_sim_stepped, ft), F(lmpc_sim_stepped, ny), F(lmpc_sim_stepped, Cm),
        F(lmpc_sim_stepped, Ddm), F(lmpc_sim_stepped, hm), F(lmpc_sim_stepped, Vtraj), F(lmpc_sim_stepped, nxo),
        F(lmpc_sim_stepped, ndm), F(lmpc_sim_stepped, Fo), F(lmpc_sim_stepped, Go), F(lmpc_sim_stepped, Gdo),
        F(lmpc_sim_stepped, fo), F(lmpc_sim_stepped, Co), F(lmpc_sim_stepped, Ddo), F(lmpc_sim_stepped, ho),
        F(lmpc_sim_stepped, Ko), F(lmpc_sim_stepped, Rtraj), F(lmpc_sim_stepped, Dtraj),
        F(lmpc_sim_stepped, Xhat_end)});
  if (out)
    for (int32_t i = 0; i < cap && i < (int32_t)o.size(); ++i) out[i] = o[i];
  return (int32_t)o.size();
}
