/*
 * bgw_jit.h -- run-time compilation of the general step kernel for ONE spec (bgw_specialize, include/bgw.h).
 *
 * The general kernel (bgw_dev.cuh, bgw_step_body) serves every sim the component API can describe; one CTA per env runs
 * at its own place in 9-16 k instructions and waits for instruction fetches most of the time (profiles/: no_instruction is
 * its top stall).  With the spec's scalars as compile-time constants the same source compiles to a third of that: every
 * dispatch on program / actor / observer / manager folds, loops get their trip counts, the shared-memory offsets become
 * immediates.  NVRTC (libnvrtc.so.12, opened with dlopen: no link-time dependency) compiles
 *
 *     #define BGW_JIT_PIN  s.H = ..; s.W = ..; ...         (every int scalar of DevSpec except E / env_offset / horizon)
 *     #define BGW_JIT_T    <threads per env>
 *     #include "bgw_dev.cuh"
 *     extern "C" __global__ void bgw_step_jit(...) { bgw_step_body<PROG, ATT>(...); }
 *
 * for sm_100a into a cubin, which is loaded through the driver API (libcuda.so.1, dlopen) into the primary context the
 * runtime API uses, and launched with cuLaunchKernel.  Host code only; included by bgw.cu.
 */
#ifndef BGW_JIT_H_
#define BGW_JIT_H_

#include <dlfcn.h>
#include <sys/stat.h>

#include <string>

namespace bgwjit {

typedef struct _nvrtcProgram *nvrtcProgram;
typedef struct CUmod_st *CUmodule;
typedef struct CUfunc_st *CUfunction;

struct Nvrtc {
    void *lib = nullptr;
    int (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
    int (*CompileProgram)(nvrtcProgram, int, const char *const *) = nullptr;
    int (*GetCUBINSize)(nvrtcProgram, size_t *) = nullptr;
    int (*GetCUBIN)(nvrtcProgram, char *) = nullptr;
    int (*GetProgramLogSize)(nvrtcProgram, size_t *) = nullptr;
    int (*GetProgramLog)(nvrtcProgram, char *) = nullptr;
    int (*DestroyProgram)(nvrtcProgram *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};

struct Driver {
    void *lib = nullptr;
    int (*ModuleLoadData)(CUmodule *, const void *) = nullptr;
    int (*ModuleGetFunction)(CUfunction *, CUmodule, const char *) = nullptr;
    int (*ModuleUnload)(CUmodule) = nullptr;
    int (*FuncSetAttribute)(CUfunction, int, int) = nullptr;
    int (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void *, void **, void **) = nullptr;
    int (*GetErrorString)(int, const char **) = nullptr;
};

template <typename F>
inline bool sym(void *lib, const char *name, F &out)
{
    out = reinterpret_cast<F>(dlsym(lib, name));
    return out != nullptr;
}

inline const char *load_nvrtc(Nvrtc &n)
{
    if (n.lib) return nullptr;
    const char *names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"};
    for (const char *nm : names) if ((n.lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL))) break;
    if (!n.lib) return "libnvrtc.so.12 not found (looked in the loader path and /usr/local/cuda/lib64)";
    if (!sym(n.lib, "nvrtcCreateProgram", n.CreateProgram) || !sym(n.lib, "nvrtcCompileProgram", n.CompileProgram) ||
        !sym(n.lib, "nvrtcGetCUBINSize", n.GetCUBINSize) || !sym(n.lib, "nvrtcGetCUBIN", n.GetCUBIN) ||
        !sym(n.lib, "nvrtcGetProgramLogSize", n.GetProgramLogSize) || !sym(n.lib, "nvrtcGetProgramLog", n.GetProgramLog) ||
        !sym(n.lib, "nvrtcDestroyProgram", n.DestroyProgram) || !sym(n.lib, "nvrtcGetErrorString", n.GetErrorString)) {
        dlclose(n.lib); n.lib = nullptr;
        return "libnvrtc lacks an entry point (nvrtcGetCUBIN needs CUDA >= 11.1)";
    }
    return nullptr;
}

inline const char *load_driver(Driver &d)
{
    if (d.lib) return nullptr;
    if (!(d.lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL))) return "libcuda.so.1 not found";
    if (!sym(d.lib, "cuModuleLoadData", d.ModuleLoadData) || !sym(d.lib, "cuModuleGetFunction", d.ModuleGetFunction) ||
        !sym(d.lib, "cuModuleUnload", d.ModuleUnload) || !sym(d.lib, "cuFuncSetAttribute", d.FuncSetAttribute) ||
        !sym(d.lib, "cuLaunchKernel", d.LaunchKernel) || !sym(d.lib, "cuGetErrorString", d.GetErrorString)) {
        dlclose(d.lib); d.lib = nullptr;
        return "libcuda.so.1 lacks an entry point";
    }
    return nullptr;
}

inline Nvrtc &nvrtc() { static Nvrtc n; return n; }
inline Driver &driver() { static Driver d; return d; }

/* directory of the shared object this code lives in (= abmarl_b200/csrc: the kernel sources sit next to libbgw.so) */
inline std::string library_dir()
{
    Dl_info info;
    static int anchor;
    if (!dladdr((void *)&anchor, &info) || !info.dli_fname) return ".";
    std::string p(info.dli_fname);
    const size_t k = p.find_last_of('/');
    return k == std::string::npos ? std::string(".") : p.substr(0, k);
}

inline bool read_file(const std::string &path, std::string &out)
{
    FILE *fp = fopen(path.c_str(), "rb");
    if (!fp) return false;
    char buf[1 << 16];
    size_t n;
    out.clear();
    while ((n = fread(buf, 1, sizeof(buf), fp)) > 0) out.append(buf, n);
    fclose(fp);
    return true;
}

inline unsigned long long fnv1a(const std::string &s, unsigned long long h = 1469598103934665603ull)
{
    for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
    return h;
}

/* the statements that overwrite the spec's scalars at the top of bgw_step_body */
inline std::string pin_statements(const DevSpec &d, bool has_init_ammo)
{
    std::string s;
    char b[96];
#define PIN(f) do { snprintf(b, sizeof(b), "s." #f "=%d;", (int)d.f); s += b; } while (0)
    PIN(H); PIN(W); PIN(HW); PIN(A); PIN(L);
    PIN(program); PIN(move_actor); PIN(attack_actor); PIN(observer); PIN(observe_self); PIN(done_mask); PIN(manager); PIN(ravel);
    PIN(no_overlap); PIN(stacked); PIN(auto_reset);
    PIN(max_enc); PIN(n_blk); PIN(obs_h); PIN(obs_w); PIN(obs_c); PIN(obs_stride); PIN(nchunks);
    PIN(a_nav); PIN(a_target); PIN(a_pacman); PIN(has_food);
    for (int k = 0; k < 10; ++k) { snprintf(b, sizeof(b), "s.a_script[%d]=%d;", k, d.a_script[k]); s += b; }
    PIN(hw_words); PIN(n_var); PIN(randomize_placement_order); PIN(tpl_error); PIN(slot_mask); PIN(mask_words); PIN(mask_batch);
    PIN(parallel_actors); PIN(act_words); PIN(ammo_offset); PIN(n_ammo); PIN(position_offset); PIN(obs_cells); PIN(n_dyn_blk);
    PIN(o_head); PIN(o_slot); PIN(o_cell); PIN(o_next); PIN(o_flags); PIN(o_enc); PIN(o_klass); PIN(o_tmp); PIN(o_racc); PIN(o_act);
    PIN(o_ragent); PIN(o_plist); PIN(o_pstate); PIN(o_avail); PIN(o_mask); PIN(o_ctr); PIN(o_csum); PIN(smem_bytes);
#undef PIN
    if (!d.static_mask) s += "s.static_mask=nullptr;";
    if (!has_init_ammo) s += "s.init_ammo=nullptr;";
    return s;
}

struct Kernel {
    CUmodule module = nullptr;
    CUfunction function = nullptr;
};

/* Compile (or fetch from cache_dir) and load the kernel.  Returns nullptr on success, else a message (static or in `msg`). */
inline const char *build(const DevSpec &d, bool has_init_ammo, int threads, const char *cache_dir, Kernel &out, std::string &msg)
{
    if (const char *e = load_driver(driver())) return e;
    const std::string dir = library_dir(), inc = dir + "/../../include";
    std::string dev, hdr, phl, sti;
    if (!read_file(dir + "/bgw_dev.cuh", dev) || !read_file(inc + "/bgw.h", hdr) || !read_file(inc + "/bgw_philox.h", phl) ||
        !read_file(inc + "/bgw_stdint.h", sti)) {
        msg = "the kernel sources (bgw_dev.cuh, include/bgw*.h) are not next to the library in " + dir;
        return msg.c_str();
    }
    char head[1024];
    snprintf(head, sizeof(head), "#define BGW_JIT_T %d\n#define BGW_JIT_PIN ", threads);
    std::string src = head + pin_statements(d, has_init_ammo) + "\n#include \"bgw_dev.cuh\"\n";
    snprintf(head, sizeof(head),
             "extern \"C\" __global__ void __launch_bounds__(%d) bgw_step_jit(const DevSpec s_in, const BgwState st, const uint32_t *actions, "
             "const int16_t *order, int8_t *obs, float *reward, uint8_t *done, uint8_t *all_done)\n"
             "{ bgw_step_body<%d, %d>(s_in, st, actions, order, obs, reward, done, all_done); }\n",
             threads, d.program, d.attack_actor);
    src += head;
    const unsigned long long key = fnv1a(sti, fnv1a(phl, fnv1a(hdr, fnv1a(dev, fnv1a(src)))));
    std::string cubin, cache_path;
    if (cache_dir && *cache_dir) {
        char nm[64];
        snprintf(nm, sizeof(nm), "/bgw_jit_%016llx.cubin", key);
        cache_path = std::string(cache_dir) + nm;
        read_file(cache_path, cubin);
    }
    if (cubin.empty()) {
        if (const char *e = load_nvrtc(nvrtc())) return e;
        Nvrtc &n = nvrtc();
        nvrtcProgram prog = nullptr;
        int rc = n.CreateProgram(&prog, src.c_str(), "bgw_jit.cu", 0, nullptr, nullptr);
        if (rc) { msg = std::string("nvrtcCreateProgram: ") + n.GetErrorString(rc); return msg.c_str(); }
        const std::string i1 = "-I" + dir, i2 = "-I" + inc;
        const char *opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "--fmad=false", i1.c_str(), i2.c_str(), "-I/usr/local/cuda/include", "-lineinfo"};
        rc = n.CompileProgram(prog, (int)(sizeof(opts) / sizeof(opts[0])), opts);
        if (rc) {
            size_t ln = 0;
            n.GetProgramLogSize(prog, &ln);
            std::string log(ln, '\0');
            if (ln) n.GetProgramLog(prog, &log[0]);
            msg = std::string("nvrtcCompileProgram: ") + n.GetErrorString(rc) + "\n" + log.substr(0, 1500);
            n.DestroyProgram(&prog);
            return msg.c_str();
        }
        size_t cn = 0;
        n.GetCUBINSize(prog, &cn);
        cubin.resize(cn);
        n.GetCUBIN(prog, &cubin[0]);
        n.DestroyProgram(&prog);
        if (!cache_path.empty()) {                    /* best effort: written under a temporary name, then renamed */
            mkdir(cache_dir, 0755);
            const std::string tmp = cache_path + ".tmp";
            if (FILE *fp = fopen(tmp.c_str(), "wb")) {
                const bool ok = fwrite(cubin.data(), 1, cubin.size(), fp) == cubin.size();
                fclose(fp);
                if (ok) rename(tmp.c_str(), cache_path.c_str()); else remove(tmp.c_str());
            }
        }
    }
    Driver &c = driver();
    const char *es = nullptr;
    int rc = c.ModuleLoadData(&out.module, cubin.data());
    if (rc) { c.GetErrorString(rc, &es); msg = std::string("cuModuleLoadData: ") + (es ? es : "?"); return msg.c_str(); }
    rc = c.ModuleGetFunction(&out.function, out.module, "bgw_step_jit");
    if (!rc) rc = c.FuncSetAttribute(out.function, 8 /* CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES */, d.smem_bytes);
    if (rc) {
        c.GetErrorString(rc, &es);
        msg = std::string("loading bgw_step_jit: ") + (es ? es : "?");
        c.ModuleUnload(out.module);
        out = Kernel();
        return msg.c_str();
    }
    return nullptr;
}

}   // namespace bgwjit

#endif /* BGW_JIT_H_ */
