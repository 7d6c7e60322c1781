/*
 * multi_api.cpp -- one population spread over the GPUs of a box, inside the library (north star (4): "particles shard
 * trivially across the 8 GPUs of one box: each GPU holds a full ephemeris copy and its own particle slice, with no
 * NCCL on the hot path, only a host gather of outputs").
 *
 * Host code on top of the per-device batch calls of assist_gpu.h: one assist_gpu_batch per device, one host thread
 * per device for every call that launches work, scatter / gather of contiguous host slices around them.  The
 * reference has no counterpart (it is single-threaded, one simulation per process: SURVEY section 2.2); a caller of
 * the reference who loops over simulations gets the same semantics from ONE call here.
 *
 * The deal: systems are sorted by their expected step count (ab_gpu_cost_order_host, the same estimate the work queue
 * uses) and dealt out round-robin in that order, so every device gets the same mix of long and short systems whatever
 * the order of the caller's population is.  Results do not depend on the deal: systems never interact.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <thread>
#include <vector>

#include "assist_gpu.h"
#include "host_internal.h"

struct assist_gpu_multi {
    int n, K, n_var, n_dev;
    std::vector<int> device;                 /* [n_dev] */
    std::vector<assist_gpu_batch*> batch;    /* [n_dev] */
    std::vector<int> count;                  /* [n_dev] systems on the device */
    std::vector<int> dev_of, slot_of;        /* [n] where system i lives */
    std::vector<std::vector<int>> members;   /* [n_dev][count] systems of the device, in slot order */
    bool dealt;
};

static thread_local char g_merr[512] = "";
extern "C" const char* assist_gpu_multi_last_error(void) { return g_merr; }

static int merr(int code, const char* msg) {
    snprintf(g_merr, sizeof(g_merr), "%s", msg);
    return code;
}

/* run f(d) for every device on its own host thread; first non-zero return code wins */
template <class F>
static int on_every_device(assist_gpu_multi* m, F f) {
    std::vector<int> rc((size_t)m->n_dev, 0);
    std::vector<std::string> msg((size_t)m->n_dev);
    std::vector<std::thread> th;
    for (int d = 0; d < m->n_dev; d++) {
        th.emplace_back([&, d]() {
            rc[(size_t)d] = assist_gpu_set_device(m->device[(size_t)d]);
            if (!rc[(size_t)d]) rc[(size_t)d] = f(d);
            if (rc[(size_t)d]) msg[(size_t)d] = assist_gpu_last_error();
        });
    }
    for (auto& t : th) t.join();
    for (int d = 0; d < m->n_dev; d++)
        if (rc[(size_t)d]) {
            char buf[480];
            snprintf(buf, sizeof(buf), "device %d: %s", m->device[(size_t)d], msg[(size_t)d].c_str());
            return merr(rc[(size_t)d], buf);
        }
    return 0;
}

extern "C" assist_gpu_multi* assist_gpu_multi_create(const struct assist_ephem* ephem, int n_sys, int n_var,
                                                     const int* devices, int n_devices) {
    if (!ephem || n_sys < 1 || n_var < 0) { merr(ASSIST_GPU_ERR_ARG, "bad argument"); return nullptr; }
    const int have = assist_gpu_device_count();
    if (have < 1) { merr(ASSIST_GPU_ERR_NO_DEVICE, "no CUDA device available: assist-b200 has no CPU compute path"); return nullptr; }
    assist_gpu_multi* m = new assist_gpu_multi();
    m->n = n_sys; m->n_var = n_var; m->K = 1 + n_var; m->dealt = false;
    if (devices && n_devices > 0) m->device.assign(devices, devices + n_devices);
    else for (int d = 0; d < have; d++) m->device.push_back(d);
    if ((int)m->device.size() > n_sys) m->device.resize((size_t)n_sys);
    m->n_dev = (int)m->device.size();
    m->batch.assign((size_t)m->n_dev, nullptr);
    m->count.assign((size_t)m->n_dev, 0);
    for (int d = 0; d < m->n_dev; d++) m->count[(size_t)d] = (n_sys - d + m->n_dev - 1) / m->n_dev;
    m->dev_of.assign((size_t)n_sys, 0); m->slot_of.assign((size_t)n_sys, 0);
    m->members.assign((size_t)m->n_dev, std::vector<int>());
    /* sequentially: the ephemeris upload of a device is not re-entrant per ephemeris handle */
    for (int d = 0; d < m->n_dev; d++) {
        int rc = assist_gpu_set_device(m->device[(size_t)d]);
        if (!rc) rc = assist_gpu_ephem_upload(ephem);
        if (!rc) {
            m->batch[(size_t)d] = assist_gpu_batch_create(ephem, m->count[(size_t)d], n_var, ASSIST_GPU_PER_PARTICLE);
            if (!m->batch[(size_t)d]) rc = ASSIST_GPU_ERR_CUDA;
        }
        if (rc) {
            char buf[480];
            snprintf(buf, sizeof(buf), "device %d: %s", m->device[(size_t)d], assist_gpu_last_error());
            merr(rc, buf);
            assist_gpu_multi_free(m);
            return nullptr;
        }
    }
    return m;
}

extern "C" void assist_gpu_multi_free(assist_gpu_multi* m) {
    if (!m) return;
    for (int d = 0; d < m->n_dev; d++)
        if (m->batch[(size_t)d]) { assist_gpu_set_device(m->device[(size_t)d]); assist_gpu_batch_free(m->batch[(size_t)d]); }
    delete m;
}

extern "C" int assist_gpu_multi_device_count(const assist_gpu_multi* m) { return m ? m->n_dev : 0; }

extern "C" int assist_gpu_multi_set_options(assist_gpu_multi* m, const struct assist_gpu_options* opt) {
    if (!m || !opt) return merr(ASSIST_GPU_ERR_ARG, "NULL argument");
    for (int d = 0; d < m->n_dev; d++) {
        int rc = assist_gpu_batch_set_options(m->batch[(size_t)d], opt);
        if (rc) return merr(rc, assist_gpu_last_error());
    }
    return 0;
}

extern "C" int assist_gpu_multi_set_state(assist_gpu_multi* m, double t0, double dt0, const double* state, const double* params) {
    if (!m || !state) return merr(ASSIST_GPU_ERR_ARG, "NULL argument");
    const size_t n = (size_t)m->n, K = (size_t)m->K;
    /* the deal: round-robin over the systems in the order of their expected step counts, longest first */
    std::vector<int> order(n);
    ab_gpu_cost_order_host(state, m->n, m->K, order.data());
    for (int d = 0; d < m->n_dev; d++) m->members[(size_t)d].clear();
    for (size_t k = 0; k < n; k++) {
        const int d = (int)(k % (size_t)m->n_dev);
        const int sys = order[k];
        m->dev_of[(size_t)sys] = d;
        m->slot_of[(size_t)sys] = (int)m->members[(size_t)d].size();
        m->members[(size_t)d].push_back(sys);
    }
    m->dealt = true;
    return on_every_device(m, [&](int d) {
        const std::vector<int>& mem = m->members[(size_t)d];
        std::vector<double> st(mem.size() * K * 6), pr(params ? mem.size() * K * 3 : 0);
        for (size_t s = 0; s < mem.size(); s++) {
            memcpy(&st[s * K * 6], state + (size_t)mem[s] * K * 6, sizeof(double) * K * 6);
            if (params) memcpy(&pr[s * K * 3], params + (size_t)mem[s] * K * 3, sizeof(double) * K * 3);
        }
        return assist_gpu_batch_set_state(m->batch[(size_t)d], t0, dt0, st.data(), params ? pr.data() : nullptr, nullptr);
    });
}

extern "C" int assist_gpu_multi_integrate(assist_gpu_multi* m, double t_end, int exact_finish_time) {
    if (!m) return merr(ASSIST_GPU_ERR_ARG, "NULL argument");
    if (!m->dealt) return merr(ASSIST_GPU_ERR_ARG, "assist_gpu_multi_set_state has not been called");
    return on_every_device(m, [&](int d) { return assist_gpu_batch_integrate(m->batch[(size_t)d], t_end, exact_finish_time, 0); });
}

extern "C" int assist_gpu_multi_integrate_or_interpolate(assist_gpu_multi* m, const double* times, int n_times, double* out) {
    if (!m || !times || !out || n_times < 1) return merr(ASSIST_GPU_ERR_ARG, "bad argument");
    if (!m->dealt) return merr(ASSIST_GPU_ERR_ARG, "assist_gpu_multi_set_state has not been called");
    const size_t n = (size_t)m->n, row = (size_t)m->K * 6;
    return on_every_device(m, [&](int d) {
        const std::vector<int>& mem = m->members[(size_t)d];
        std::vector<double> o((size_t)n_times * mem.size() * row);
        int rc = assist_gpu_batch_integrate_or_interpolate(m->batch[(size_t)d], times, n_times, o.data());
        if (rc) return rc;
        for (size_t e = 0; e < (size_t)n_times; e++)
            for (size_t s = 0; s < mem.size(); s++)
                memcpy(out + (e * n + (size_t)mem[s]) * row, &o[(e * mem.size() + s) * row], sizeof(double) * row);
        return 0;
    });
}

/* the host gather of outputs: every array in the order of the caller's population */
extern "C" int assist_gpu_multi_get_state(assist_gpu_multi* m, double* state, double* t, double* dt, double* dt_last_done, int* status) {
    if (!m) return merr(ASSIST_GPU_ERR_ARG, "NULL argument");
    if (!m->dealt) return merr(ASSIST_GPU_ERR_ARG, "assist_gpu_multi_set_state has not been called");
    const size_t row = (size_t)m->K * 6;
    return on_every_device(m, [&](int d) {
        const std::vector<int>& mem = m->members[(size_t)d];
        std::vector<double> st(state ? mem.size() * row : 0), tt(t ? mem.size() : 0), dd(dt ? mem.size() : 0), dl(dt_last_done ? mem.size() : 0);
        std::vector<int> ss(status ? mem.size() : 0);
        int rc = assist_gpu_batch_get_state(m->batch[(size_t)d], state ? st.data() : nullptr, nullptr, t ? tt.data() : nullptr,
                                            dt ? dd.data() : nullptr, dt_last_done ? dl.data() : nullptr, status ? ss.data() : nullptr);
        if (rc) return rc;
        for (size_t s = 0; s < mem.size(); s++) {
            const size_t i = (size_t)mem[s];
            if (state) memcpy(state + i * row, &st[s * row], sizeof(double) * row);
            if (t) t[i] = tt[s];
            if (dt) dt[i] = dd[s];
            if (dt_last_done) dt_last_done[i] = dl[s];
            if (status) status[i] = ss[s];
        }
        return 0;
    });
}

extern "C" int assist_gpu_multi_get_counters(assist_gpu_multi* m, unsigned long long* steps, unsigned long long* rejected,
                                             unsigned long long* iters, unsigned long long* evals) {
    if (!m) return merr(ASSIST_GPU_ERR_ARG, "NULL argument");
    if (!m->dealt) return merr(ASSIST_GPU_ERR_ARG, "assist_gpu_multi_set_state has not been called");
    return on_every_device(m, [&](int d) {
        const std::vector<int>& mem = m->members[(size_t)d];
        std::vector<unsigned long long> a(mem.size()), b(mem.size()), c(mem.size()), e(mem.size());
        int rc = assist_gpu_batch_get_counters(m->batch[(size_t)d], a.data(), b.data(), c.data(), e.data());
        if (rc) return rc;
        for (size_t s = 0; s < mem.size(); s++) {
            const size_t i = (size_t)mem[s];
            if (steps) steps[i] = a[s];
            if (rejected) rejected[i] = b[s];
            if (iters) iters[i] = c[s];
            if (evals) evals[i] = e[s];
        }
        return 0;
    });
}

/* sums over the devices; last_kernel_ms = the slowest device (the devices run side by side) */
extern "C" int assist_gpu_multi_get_stats(assist_gpu_multi* m, struct assist_gpu_stats* stats, double* kernel_ms_per_device) {
    if (!m || !stats) return merr(ASSIST_GPU_ERR_ARG, "NULL argument");
    memset(stats, 0, sizeof(*stats));
    for (int d = 0; d < m->n_dev; d++) {
        struct assist_gpu_stats s;
        int rc = assist_gpu_batch_get_stats(m->batch[(size_t)d], &s);
        if (rc) return merr(rc, assist_gpu_last_error());
        stats->steps += s.steps; stats->steps_rejected += s.steps_rejected; stats->pc_iterations += s.pc_iterations;
        stats->force_evals += s.force_evals; stats->kernel_launches += s.kernel_launches;
        if (s.last_kernel_ms > stats->last_kernel_ms) stats->last_kernel_ms = s.last_kernel_ms;
        if (kernel_ms_per_device) kernel_ms_per_device[d] = s.last_kernel_ms;
    }
    return 0;
}
