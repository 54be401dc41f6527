// Simulation mode: the reference's SECOND caller of the step -- the UI loop
// (backend/simulation/controller.py:223-269 -> backend/rocket/controls.py:27-113) -- and its wire
// format (backend/protocol.py), as a mode of the same step math.
// Differences from the gym env: the action is clipped before BOTH physics and reward
// (controls.py:72-75); there is no max-step truncation and no auto-reset; a rocket freezes after
// touchdown and is reported inactive from the next tick on; nothing is clipped for output.  Every
// tick emits one 64-byte BinaryProtocol record per rocket straight from the state planes.
#pragma once
#include "rl_kernels.cuh"

namespace rl {

// meta bits used only by this mode (the gym step clears them): landed flag and final landing code
constexpr uint32_t kMetaLanded = 0x40000000u;
constexpr int kMetaCodeShift = 27;                 // bits [27, 30)

template <class PP>
__global__ void __launch_bounds__(kCtaThreads) sim_step_kernel(const __grid_constant__ DevParams Prt, StatePlanes st,
                                                               const float2* __restrict__ actions,
                                                               float4* __restrict__ telemetry,   // [N][4] float4 = 16 f32
                                                               float* __restrict__ reward, uint8_t* __restrict__ done,
                                                               uint32_t n) {
    const auto& P = ParamSel<PP>::get(Prt);
    for (uint32_t i = blockIdx.x * kCtaThreads + threadIdx.x; i < n; i += gridDim.x * kCtaThreads) {
        Rocket r;
        load_rocket(st, i, r);
        float4* rec = telemetry + 4 * (size_t)i;
        if (r.meta & kMetaLanded) {
            // controller.py:233-240 skips it; protocol.py:63-76 encodes an inactive rocket:
            // zeros, reward NaN, no action, the remembered landing code, is_active 0
            const float code = (float)((r.meta >> kMetaCodeShift) & 7u);
            rec[0] = make_float4(0.f, 0.f, 0.f, 0.f);
            rec[1] = make_float4(0.f, 0.f, 0.f, 0.f);
            rec[2] = make_float4(0.f, 0.f, 0.f, __int_as_float(0x7fc00000));
            rec[3] = make_float4(0.f, 0.f, code, 0.f);
            if (reward) reward[i] = __int_as_float(0x7fc00000);
            if (done) done[i] = 1;
            continue;
        }
        const float2 act = actions[i];
        const float th = clipf(act.x, 0.0f, 1.0f), cg = clipf(act.y, -1.0f, 1.0f);     // controls.py:72-73
        StepResult o;
        rocket_step<PP, true>(P, r, th, cg, o);           // the reward sees the CLIPPED action here; positions clamp
                                                          // at the edge of the fixed-point grid (no bounds check in this mode)
        const float vx = o.vx, vy = o.vy, om = o.om;      // (delta / dt of the state just written)
        const bool landed = o.terminated;                 // the truncation flag is ignored (controls.py:89)
        if (landed) r.meta |= kMetaLanded | ((uint32_t)o.landing << kMetaCodeShift);
        store_rocket(st, i, r, false);
        // protocol.py:78-96; the echoed action is the one just applied, zeros on the touchdown tick
        rec[0] = make_float4((float)r.x * P.pos_unit, (float)r.y * P.pos_unit, vx, vy);
        rec[1] = make_float4(o.ax, o.ay, angle_deg(r.ang), om);
        rec[2] = make_float4(o.alpha, r.dry, (float)((double)r.fuel * P.fuel_unit_d), o.reward);
        rec[3] = make_float4(landed ? 0.0f : act.x, landed ? 0.0f : act.y, (float)o.landing, 1.0f);
        if (reward) reward[i] = o.reward;
        if (done) done[i] = landed ? 1 : 0;
    }
}

}  // namespace rl
