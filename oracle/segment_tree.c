/*
 * oracle/segment_tree.c -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Plain-C (fp64, glibc libm) restatement of the arithmetic of
 *   rl_coach/memories/non_episodic/prioritized_experience_replay.py
 *     SegmentTree._propagate        :63-74     -> ost_update()
 *     SegmentTree._retrieve         :76-92     -> ost_retrieve()
 *     SegmentTree.add               :102-114   -> (cursor handled by the caller, see oper_store())
 *     PER._update_priority          :188-201   -> oper_update_priorities()
 *     PER.update_priorities         :203-217   -> oper_update_priorities()
 *     PER.sample                    :219-262   -> oper_sample()
 *     PER.store                     :264-283   -> oper_store()
 * The reference walks the tree recursively in Python; the walks here are iterative but perform the same
 * floating-point operations in the same order, so results are bit-identical (checked in
 * tests/test_oracle_pinned.py against the imported reference and its known-answer tests).
 *
 * Tree layout = the reference's: implicit heap, tree[0] root, children of p at 2p+1 / 2p+2, leaves at
 * [size-1, 2*size-1), all float64.
 *
 * Build: make -C oracle   (gcc -O2 -fPIC -shared; -ffp-contract=off so no FMA is fused into the fp64 math).
 */
#include <math.h>
#include <stdint.h>

enum { OST_SUM = 0, OST_MIN = 1, OST_MAX = 2 };

static inline double ost_op(int op, double a, double b)
{
    if (op == OST_SUM) return a + b;
    /* Python's min(a, b) returns a unless b < a; max(a, b) returns a unless b > a (builtin semantics). */
    if (op == OST_MIN) return (b < a) ? b : a;
    return (b > a) ? b : a;
}

void ost_init(double *tree, int64_t size, int op)
{
    double v = (op == OST_SUM) ? 0.0 : (op == OST_MIN ? INFINITY : -INFINITY);
    for (int64_t i = 0; i < 2 * size - 1; ++i) tree[i] = v;
}

/* SegmentTree.update :116-129 + _propagate :63-74 */
void ost_update(double *tree, int64_t size, int op, int64_t leaf, double val)
{
    int64_t node = leaf + size - 1;
    tree[node] = val;
    while (node != 0) {
        int64_t parent = (node - 1) / 2;
        tree[parent] = ost_op(op, tree[2 * parent + 1], tree[2 * parent + 2]);
        node = parent;
    }
}

/* SegmentTree._retrieve :76-92 (returns the *leaf* index, i.e. node - size + 1 as in :141-142) */
int64_t ost_retrieve(const double *tree, int64_t size, double val)
{
    int64_t n = 2 * size - 1, node = 0;
    for (;;) {
        int64_t left = 2 * node + 1;
        if (left >= n) break;
        if (val <= tree[left]) {
            node = left;
        } else {
            val -= tree[left];
            node = left + 1;
        }
    }
    return node - size + 1;
}

/* PER.update_priorities :203-217 -- sequential, duplicates: last writer wins. Returns -1 on a negative error
 * (reference raises ValueError :195-196 *after* having applied the earlier entries), else 0.
 * *max_priority receives max_tree.total_value() (:201). */
int oper_update_priorities(double *sum_t, double *min_t, double *max_t, int64_t size,
                           const int64_t *idx, const double *err, int64_t n,
                           double epsilon, double alpha, double *max_priority)
{
    for (int64_t i = 0; i < n; ++i) {
        if (err[i] < 0) return -1;
        double p = err[i] + epsilon;
        double pa = pow(p, alpha);
        ost_update(sum_t, size, OST_SUM, idx[i], pa);
        ost_update(min_t, size, OST_MIN, idx[i], pa);
        ost_update(max_t, size, OST_MAX, idx[i], p);
        *max_priority = max_t[0];
    }
    return 0;
}

/* PER.store :264-283 for n consecutive transitions: leaf priority = maximal_priority (raw) into the max tree and
 * maximal_priority**alpha into sum/min; ring cursor wraps at size (:112-114). Returns the new cursor. */
int64_t oper_store(double *sum_t, double *min_t, double *max_t, int64_t size, int64_t cursor, int64_t n,
                   double maximal_priority, double alpha)
{
    double pa = pow(maximal_priority, alpha);
    for (int64_t i = 0; i < n; ++i) {
        ost_update(sum_t, size, OST_SUM, cursor, pa);
        ost_update(min_t, size, OST_MIN, cursor, pa);
        ost_update(max_t, size, OST_MAX, cursor, maximal_priority);
        cursor += 1;
        if (cursor >= size) cursor = 0;
    }
    return cursor;
}

/* PER.sample :229-253.  u[i] are the raw random.random() draws; random.uniform(a, b) = a + (b - a) * u
 * (CPython Lib/random.py).  nt = num_transitions() (the doubled count of quirk Q1). */
void oper_sample(const double *sum_t, const double *min_t, int64_t size, int64_t n, const double *u,
                 int64_t nt, double beta, int64_t *out_idx, double *out_w, double *out_val)
{
    double total = sum_t[0];
    double segment = total / (double)n;
    double min_probability = min_t[0] / total;
    double max_weight = pow(min_probability * (double)nt, -beta);
    for (int64_t i = 0; i < n; ++i) {
        double a = segment * (double)i;
        double b = segment * (double)(i + 1);
        double val = a + (b - a) * u[i];
        int64_t leaf = ost_retrieve(sum_t, size, val);
        double priority = sum_t[leaf + size - 1];
        priority /= total;
        double w = pow((double)nt * priority, -beta);
        out_idx[i] = leaf;
        out_w[i] = w / max_weight;
        if (out_val) out_val[i] = val;
    }
}
