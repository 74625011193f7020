// psb_fsg_racecheck.h -- TEST BUILDS ONLY (-DPSB_FSG_HOST_EMUL -DPSB_FSG_RACECHECK): a data-race
// detector for the block-wide phase code of psb_fsg_core.h / psb_ngs_core.h / psb_ngf_core.h.
//
// The plain host emulation runs every FSG_FOR loop to completion, which hides what only exists on
// the device: threads of one phase run concurrently and FSG_SYNC() is the only ordering.  Here every
// access to an utterance's mutable state (fsg_wp / fsg_wup arrays, fsg_int scalars) is recorded
// with (phase, thread); a phase ends at FSG_SYNC().  Reported:
//   * a location written by one thread and read or written by another in the same phase
//     (same-value double writes are tolerated: they cannot change the outcome),
//   * a plain access to a location that is updated atomically in the same phase.
// "Thread" identities are conservative: every loop index is its own thread, the leader section is a
// thread of its own, and code outside FSG_FOR / FSG_IF_LEADER is executed by ALL threads (a read
// there conflicts with any single thread's write in the phase; that is exactly the pattern
// "everyone reads a shared scalar, the leader updates it, no barrier in between").
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <cstddef>
#include <unordered_map>

namespace fsgrace {
enum { T_ALL = -3, T_LEADER = -1, T_ATOMIC = -4, T_MULTI = -5 };
struct Shadow { long w_phase = -1, r_phase = -1; int w_tid = 0, r_tid = 0; uint32_t w_val = 0; };
struct State {
    long phase = 0;
    int tid = T_ALL;
    long n_races = 0;
    bool off = false;
    std::unordered_map<const void *, Shadow> sh;
};
inline State &st() { static State s; return s; }
inline void report(const char *what, const void *p, const Shadow &s)
{
    State &g = st();
    if (++g.n_races <= 12)
        fprintf(stderr, "RACE %s at %p: phase %ld thread %d (last write phase %ld thread %d, last read phase %ld thread %d)\n", what, p,
                g.phase, g.tid, s.w_phase, s.w_tid, s.r_phase, s.r_tid);
}
inline void on_read(const void *p)
{
    State &g = st();
    if (g.off) return;
    Shadow &s = g.sh[p];
    if (s.w_phase == g.phase && (s.w_tid != g.tid || g.tid == T_ATOMIC)) report("read-after-write", p, s);
    if (s.r_phase != g.phase) { s.r_phase = g.phase; s.r_tid = g.tid; }
    else if (s.r_tid != g.tid) s.r_tid = T_MULTI;
}
inline void on_write(const void *p, uint32_t v)
{
    State &g = st();
    if (g.off) return;
    Shadow &s = g.sh[p];
    if (s.w_phase == g.phase && s.w_tid != g.tid && !(s.w_tid != T_ATOMIC && s.w_val == v)) report("write-after-write", p, s);
    if (s.r_phase == g.phase && s.r_tid != g.tid) report("write-after-read", p, s);
    s.w_phase = g.phase; s.w_tid = g.tid; s.w_val = v;
}
inline void on_atomic(const void *p)
{
    State &g = st();
    if (g.off) return;
    Shadow &s = g.sh[p];
    if (s.w_phase == g.phase && s.w_tid != T_ATOMIC) report("atomic-after-plain-write", p, s);
    if (s.r_phase == g.phase) report("atomic-after-plain-read", p, s);
    s.w_phase = g.phase; s.w_tid = T_ATOMIC;
}
inline void sync() { State &g = st(); ++g.phase; g.tid = T_ALL; }
inline bool for_cond(bool c, int i) { st().tid = c ? i : (int)T_ALL; return c; }
inline int leader_begin() { st().tid = T_LEADER; return 1; }
inline int leader_end() { st().tid = T_ALL; return 0; }

template <class T> struct Ref {
    T *p;
    operator T() const { on_read(p); return *p; }
    Ref &operator=(T v) { on_write(p, (uint32_t)v); *p = v; return *this; }
    Ref &operator=(const Ref &o) { return *this = (T)o; }
    Ref &operator+=(T v) { return *this = (T)((T) * this + v); }
    Ref &operator-=(T v) { return *this = (T)((T) * this - v); }
    Ref &operator&=(T v) { return *this = (T)((T) * this & v); }
    Ref &operator|=(T v) { return *this = (T)((T) * this | v); }
};
template <class T> struct Ptr {
    T *p;
    Ptr() : p(nullptr) {}
    Ptr(std::nullptr_t) : p(nullptr) {}
    Ptr(T *q) : p(q) {}
    template <class I> Ref<T> operator[](I i) const { return Ref<T>{p + i}; }
    template <class I> Ptr operator+(I i) const { return Ptr(p + i); }
    explicit operator bool() const { return p != nullptr; }
};
template <class T> struct Scalar {
    T v;
    operator T() const { on_read(&v); return v; }
    Scalar &operator=(T x) { on_write(&v, (uint32_t)x); v = x; return *this; }
    Scalar &operator=(const Scalar &o) { return *this = (T)o; }
    Scalar &operator+=(T x) { return *this = (T)((T) * this + x); }
    Scalar &operator-=(T x) { return *this = (T)((T) * this - x); }
    Scalar &operator*=(T x) { return *this = (T)((T) * this * x); }
    Scalar &operator^=(T x) { return *this = (T)((T) * this ^ x); }
};
inline void amax(Scalar<int> *p, int v) { on_atomic(&p->v); if (v > p->v) p->v = v; }
inline void amin(Scalar<int> *p, int v) { on_atomic(&p->v); if (v < p->v) p->v = v; }
inline void aadd(Scalar<int> *p, int v) { on_atomic(&p->v); p->v += v; }
inline void amax_at(const Ptr<int32_t> &a, size_t i, int v) { on_atomic(a.p + i); if (v > a.p[i]) a.p[i] = v; }
inline void amin_at(const Ptr<int32_t> &a, size_t i, int v) { on_atomic(a.p + i); if (v < a.p[i]) a.p[i] = v; }
inline void aadd_at(const Ptr<int32_t> &a, size_t i, int v) { on_atomic(a.p + i); a.p[i] += v; }
inline int afadd_at(const Ptr<int32_t> &a, size_t i, int v) { on_atomic(a.p + i); const int o = a.p[i]; a.p[i] = o + v; return o; }
}  // namespace fsgrace

#ifdef PSB_FSG_EMUL_REVERSE
#define FSG_FOR(i, n) for (int i = (n) - 1; fsgrace::for_cond(i >= 0, i); --i)
#else
#define FSG_FOR(i, n) for (int i = 0; fsgrace::for_cond(i < (n), i); ++i)
#endif
#define FSG_SYNC() fsgrace::sync()
#define FSG_IF_LEADER for (int l_ = fsgrace::leader_begin(); l_; l_ = fsgrace::leader_end())
#define FSG_ATOMIC_MAX(p, v) fsgrace::amax((p), (v))
#define FSG_ATOMIC_MIN(p, v) fsgrace::amin((p), (v))
#define FSG_ATOMIC_ADD(p, v) fsgrace::aadd((p), (v))
#define FSG_ATOMIC_MAX_AT(a, i, v) fsgrace::amax_at((a), (i), (v))
#define FSG_ATOMIC_MIN_AT(a, i, v) fsgrace::amin_at((a), (i), (v))
#define FSG_ATOMIC_ADD_AT(a, i, v) fsgrace::aadd_at((a), (i), (v))
#define FSG_ATOMIC_FETCH_ADD_AT(a, i, v) fsgrace::afadd_at((a), (i), (v))
#define FSG_COLLECTIVE_BEGIN() fsgrace::sync()
#define FSG_COLLECTIVE_END() fsgrace::sync()
#define FSG_RAW(a) ((a).p)
typedef fsgrace::Ptr<int32_t> fsg_wp;
typedef fsgrace::Ptr<uint32_t> fsg_wup;
typedef fsgrace::Scalar<int> fsg_int;
typedef fsgrace::Scalar<float> fsg_float;
typedef fsgrace::Scalar<long long> fsg_ll;
