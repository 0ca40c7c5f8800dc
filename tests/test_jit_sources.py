"""The kernel sources stay compilable at run time (bgw_specialize, abmarl_b200/csrc/bgw_jit.h): NVRTC has no host headers
and rejects unannotated host functions, so every header the device code includes must guard them under __CUDACC_RTC__.
NVRTC cross-compiles for sm_100a without a GPU (the toolkit's libnvrtc through cuda-python)."""
import os

import pytest

nvrtc = pytest.importorskip('cuda.bindings.nvrtc')

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, 'abmarl_b200', 'csrc')

GENERAL = '''
#define BGW_JIT_T 32
#define BGW_JIT_PIN s.H=6;s.W=6;s.HW=36;s.A=18;s.L=18;s.move_actor=1;s.observer=0;s.observe_self=1;s.manager=0;s.ravel=0;s.stacked=0;s.n_blk=0;s.static_mask=nullptr;
#include "bgw_dev.cuh"
extern "C" __global__ void __launch_bounds__(32) bgw_step_jit(const DevSpec s_in, const BgwState st, const uint32_t *actions,
    const int16_t *order, int8_t *obs, float *reward, uint8_t *done, uint8_t *all_done)
{ bgw_step_body<0, 2>(s_in, st, actions, order, obs, reward, done, all_done); }
'''

FAST = '''
#include "bgw_dev.cuh"
#include "bgw_fast.cuh"
struct FastStaticJit {
    static constexpr bool is_static = true;
    static constexpr int A = 20, L = 20, H = 6, W = 7, P = 4, PL = 4, PW = 16, PH = 14, obs_stride = 96, nchunks = 6,
        obs_h = 9, view = -1, move_actor = 1, ravel = 0, observe_self = 1, done_mask = 4, max_enc = 3, simd_ok = 1,
        async_ok = 0, slots = 32, T = 32, att = -1, identity = 1, can_mix = 1, acc_lt1 = 1, rpo = 0, LB_T = 32, LB_N = 28;
};
extern "C" __global__ void __launch_bounds__(32, 28) bgw_step_fast_jit(const DevSpec s_in, const FastSpec f_in, const BgwState st,
    const uint32_t *actions, uint32_t *sampled, const int16_t *order, int8_t *obs, float *reward, uint8_t *done, uint8_t *all_done)
{ bgw_step_fast_body<FastStaticJit, uint8_t>(s_in, f_in, st, actions, sampled, order, obs, reward, done, all_done); }
'''


@pytest.mark.parametrize('name,src,extra', [('bgw_step_jit', GENERAL, []), ('bgw_step_fast_jit', FAST, []),
                                            ('bgw_step_fast_jit', FAST, [b'-DBGW_NO_ST256'])])
def test_kernel_sources_compile_under_nvrtc(name, src, extra):
    err, prog = nvrtc.nvrtcCreateProgram(src.encode(), b'bgw_jit.cu', 0, [], [])
    assert int(err) == 0
    opts = [b'--gpu-architecture=sm_100a', b'-std=c++17', b'--fmad=false', b'-I' + CSRC.encode(),
            b'-I' + os.path.join(ROOT, 'include').encode(), b'-I/usr/local/cuda/include'] + extra
    err, = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    _, n = nvrtc.nvrtcGetProgramLogSize(prog)
    log = b' ' * n
    nvrtc.nvrtcGetProgramLog(prog, log)
    assert int(err) == 0, log.decode(errors='replace')[:3000]
    _, size = nvrtc.nvrtcGetCUBINSize(prog)
    assert size > 10000
    cubin = b' ' * size
    nvrtc.nvrtcGetCUBIN(prog, cubin)
    assert name.encode() in cubin
