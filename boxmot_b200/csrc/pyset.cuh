// pyset.cuh -- iteration order of CPython's `set` for small non-negative ints, as far as StrongSORT observes it.
//
// matching_cascade returns `list(set(track_indices) - set(k for k, _ in matches))`
// (trackers/bbox/strongsort/sort/linear_assignment.py:108) and Tracker._match feeds that list, in that order, into
// the rows of the IoU stage (sort/tracker.py:139-153).  Row order decides scipy's choice among tied solutions and
// the order of the unmatched detections, hence ids; a set of small ints is NOT iterated in ascending order once a
// key exceeds the table size (`list({17, 3}) == [17, 3]`).  This restates the parts of Objects/setobject.c
// (CPython 3.7 - 3.13: open addressing, LINEAR_PROBES = 9, perturb shift 5, growth to used*4 once fill*5 >=
// mask*3; set_difference's two strategies; set_copy via set_merge) that determine that order.  hash(k) == k.
//
// Tables hold key+1 (0 = never used, -1 = dummy left by a discard).  Everything here runs on ONE thread; shortcuts
// skip the emulation whenever the table provably maps every key to its own slot (then iteration is ascending).
#pragma once
#include "tracker_core.cuh"

namespace bmb {

struct PySetTab {
    int* tab;     // current table
    int* spare;   // second buffer for resizes
    int mask, fill, used;
};

BMB_FN void pyset_init(PySetTab& t, int* a, int* b) {
    t.tab = a; t.spare = b; t.mask = 7; t.fill = 0; t.used = 0;
    for (int i = 0; i < 8; ++i) a[i] = 0;
}

BMB_FN void pyset_insert_clean(int* tab, int mask, int key) {
    unsigned perturb = (unsigned)key;
    int i = key & mask;
    while (true) {
        const int probes = (i + 9 <= mask) ? 9 : 0;
        for (int e = i; e <= i + probes; ++e)
            if (tab[e] == 0) { tab[e] = key + 1; return; }
        perturb >>= 5;
        i = (int)(((unsigned)i * 5u + 1u + perturb) & (unsigned)mask);
    }
}

BMB_FN int pyset_size_for(int minused) {
    int n = 8;
    while (n <= minused) n <<= 1;
    return n;
}

BMB_FN void pyset_resize(PySetTab& t, int minused) {
    const int newsize = pyset_size_for(minused);
    int* nt = t.spare;
    for (int i = 0; i < newsize; ++i) nt[i] = 0;
    for (int i = 0; i <= t.mask; ++i)
        if (t.tab[i] > 0) pyset_insert_clean(nt, newsize - 1, t.tab[i] - 1);
    t.spare = t.tab;
    t.tab = nt;
    t.mask = newsize - 1;
    t.fill = t.used;
}

// set_add_entry for a key known to be absent or present (duplicates are ignored)
BMB_FN void pyset_add(PySetTab& t, int key) {
    const int mask = t.mask;
    unsigned perturb = (unsigned)key;
    int i = key & mask, freeslot = -1, found = -1;
    while (found < 0) {
        const int probes = (i + 9 <= mask) ? 9 : 0;
        for (int e = i; e <= i + probes; ++e) {
            const int v = t.tab[e];
            if (v == 0) { found = e; break; }
            if (v == key + 1) return;
            if (v < 0 && freeslot < 0) freeslot = e;
        }
        if (found >= 0) break;
        perturb >>= 5;
        i = (int)(((unsigned)i * 5u + 1u + perturb) & (unsigned)mask);
    }
    if (freeslot >= 0) { t.tab[freeslot] = key + 1; t.used += 1; return; }
    t.tab[found] = key + 1;
    t.fill += 1;
    t.used += 1;
    if ((long long)t.fill * 5 < (long long)mask * 3) return;
    pyset_resize(t, t.used > 50000 ? t.used * 2 : t.used * 4);
}

// mask of a set grown from empty by `n` distinct insertions (depends on the count only)
BMB_FN int pyset_mask_after(int n) {
    int mask = 7;
    while (true) {
        const int trig = (mask * 3 + 4) / 5;  // first fill with fill*5 >= mask*3: the insertion that resizes
        if (trig > n) break;
        mask = pyset_size_for(trig > 50000 ? trig * 2 : trig * 4) - 1;
    }
    return mask;
}

BMB_FN int pyset_items(const PySetTab& t, int* out) {
    int n = 0;
    for (int i = 0; i <= t.mask; ++i)
        if (t.tab[i] > 0) out[n++] = t.tab[i] - 1;
    return n;
}

// True when list(set(a) - set(b)) is simply the ascending filter of a: na = len(a), amax = max(a), nb = len(b),
// umax = largest key of a that is not in b (-1 when there is none).
BMB_FN bool pyset_difference_is_ascending(int na, int amax, int nb, int umax) {
    if (na == 0 || umax < 0) return true;
    if ((na >> 2) > nb) {
        int cmask = 7;
        if ((long long)na * 5 >= 7 * 3) cmask = pyset_size_for(na * 2) - 1;
        return amax <= cmask;
    }
    return umax <= pyset_mask_after(na - nb);
}

// list(set(a) - set(b)) where a[0..na) is ascending and b is the subset of a flagged by is_b(key) (nb elements).
// buf0..buf3 hold at least 8*na + 16 ints each; `order` (>= na ints) is scratch for the iteration order of
// set(a).  Returns the number of keys written to `out`.
template <typename IsB>
BMB_FN int pyset_difference(const int* a, int na, int nb, IsB is_b, int* out, int* order, int* buf0, int* buf1,
                            int* buf2, int* buf3) {
    if (na == 0) return 0;
    const int amax = a[na - 1];
    // ---- iteration order of so = set(a) ----
    const int so_mask = pyset_mask_after(na);
    const bool so_identity = amax <= so_mask;
    const int* ord = a;
    if (!so_identity) {
        PySetTab so;
        pyset_init(so, buf0, buf1);
        for (int k = 0; k < na; ++k) pyset_add(so, a[k]);
        pyset_items(so, order);
        ord = order;
    }
    int n = 0;
    if ((na >> 2) > nb) {
        // set_copy_and_difference: copy (set_merge into an empty set), then discard b's keys (dummies keep order)
        int cmask = 7;
        if ((long long)na * 5 >= 7 * 3) cmask = pyset_size_for(na * 2) - 1;
        if (amax <= cmask) {
            for (int k = 0; k < na; ++k) if (!is_b(a[k])) out[n++] = a[k];       // identity mapped copy
        } else if (cmask == so_mask) {
            for (int k = 0; k < na; ++k) if (!is_b(ord[k])) out[n++] = ord[k];      // same slots as so
        } else {
            int* ct = buf2;
            for (int i = 0; i <= cmask; ++i) ct[i] = 0;
            for (int k = 0; k < na; ++k) pyset_insert_clean(ct, cmask, ord[k]);
            for (int i = 0; i <= cmask; ++i)
                if (ct[i] > 0 && !is_b(ct[i] - 1)) out[n++] = ct[i] - 1;
        }
        return n;
    }
    // general strategy: walk so, add the keys that are not in b to a fresh set
    const int nu = na - nb;
    int umax = -1;
    for (int k = na - 1; k >= 0; --k) if (!is_b(a[k])) { umax = a[k]; break; }
    if (umax < 0) return 0;
    if (umax <= pyset_mask_after(nu)) {
        for (int k = 0; k < na; ++k) if (!is_b(a[k])) out[n++] = a[k];
        return n;
    }
    PySetTab r;
    pyset_init(r, buf2, buf3);
    for (int k = 0; k < na; ++k) if (!is_b(ord[k])) pyset_add(r, ord[k]);
    return pyset_items(r, out);
}

}  // namespace bmb
