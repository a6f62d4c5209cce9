// Order-exact replay of cv::KeyPointsFilter::retainBest.
//
// cv2's ORB keeps, per pyramid level, the n best keypoints with
//     std::nth_element(begin, begin+n-1, end, response-greater); std::partition(begin+n, end, response >= boundary)
// and the ORDER those two libstdc++ algorithms leave behind decides descriptor row order, hence match indices, hence
// the positions cv2's RANSAC samples.  This header re-implements the behaviour of those two
// algorithms (introselect: median-of-3 to first, unguarded Hoare partition, depth limit 2*floor(log2 n) with
// heap-select fallback, final insertion sort on <= 3 elements; bidirectional partition) over an accessor, so that the
// same control flow runs (a) on one warp of the GPU with the two scan loops done 32 lanes at a time, and (b) on the
// host inside tests/hostsim where it is checked against the real std:: calls.
//
// Reference call site whose output order this reproduces: cv2 ORB::detectAndCompute via
// /root/reference/scripts/visual_odometry_v3.py:373.
#pragma once
#include "mathcore.cuh"

namespace dvo {

// Accessor contract:
//   typedef ... Item;
//   Item get(int i); void set(int i, Item v);
//   static bool gt(Item a, Item b)            -- comp(a, b): a's response > b's response
//   int scan_up(int first, Item pivot)        -- smallest i >= first with !gt(a[i], pivot)   (unguarded: one exists)
//   int scan_down(int last, Item pivot)       -- largest  i <= last  with !gt(pivot, a[i])   (unguarded)
//   int scan_up_pred(int first, int last, Item b)   -- smallest i in [first,last) with !(a[i] >= b), else last
//   int scan_down_pred(int first, int last, Item b) -- largest i in (first,last] ... see partition_ge

template <class Acc>
DVO_HDN void sel_swap(Acc& a, int i, int j) {
    typename Acc::Item x = a.get(i), y = a.get(j);
    a.set(i, y);
    a.set(j, x);
}

template <class Acc>
DVO_HDN void move_median_to_first(Acc& a, int result, int ia, int ib, int ic) {
    typename Acc::Item A = a.get(ia), B = a.get(ib), C = a.get(ic);
    if (Acc::gt(A, B)) {
        if (Acc::gt(B, C)) sel_swap(a, result, ib);
        else if (Acc::gt(A, C)) sel_swap(a, result, ic);
        else sel_swap(a, result, ia);
    } else if (Acc::gt(A, C)) sel_swap(a, result, ia);
    else if (Acc::gt(B, C)) sel_swap(a, result, ic);
    else sel_swap(a, result, ib);
}

template <class Acc>
DVO_HDN int unguarded_partition(Acc& a, int first, int last, int pivot_idx) {
    typename Acc::Item pivot = a.get(pivot_idx);
    while (true) {
        first = a.scan_up(first, pivot);
        --last;
        last = a.scan_down(last, pivot);
        if (!(first < last)) return first;
        sel_swap(a, first, last);
        ++first;
    }
}

template <class Acc>
DVO_HDN void adjust_heap(Acc& a, int first, int hole, int len, typename Acc::Item value) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (Acc::gt(a.get(first + child), a.get(first + child - 1))) child--;
        a.set(first + hole, a.get(first + child));
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        a.set(first + hole, a.get(first + child - 1));
        hole = child - 1;
    }
    // push_heap
    int parent = (hole - 1) / 2;
    while (hole > top && Acc::gt(a.get(first + parent), value)) {
        a.set(first + hole, a.get(first + parent));
        hole = parent;
        parent = (hole - 1) / 2;
    }
    a.set(first + hole, value);
}

template <class Acc>
DVO_HDN void heap_select(Acc& a, int first, int middle, int last) {
    int len = middle - first;
    if (len >= 2) {
        int parent = (len - 2) / 2;
        while (true) {
            typename Acc::Item v = a.get(first + parent);
            adjust_heap(a, first, parent, len, v);
            if (parent == 0) break;
            parent--;
        }
    }
    for (int i = middle; i < last; ++i) {
        if (Acc::gt(a.get(i), a.get(first))) {
            typename Acc::Item v = a.get(i);
            a.set(i, a.get(first));
            adjust_heap(a, first, 0, len, v);
        }
    }
}

template <class Acc>
DVO_HDN void insertion_sort(Acc& a, int first, int last) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        typename Acc::Item v = a.get(i);
        if (Acc::gt(v, a.get(first))) {
            for (int k = i; k > first; --k) a.set(k, a.get(k - 1));
            a.set(first, v);
        } else {
            int pos = i, next = i - 1;
            while (Acc::gt(v, a.get(next))) {
                a.set(pos, a.get(next));
                pos = next;
                --next;
            }
            a.set(pos, v);
        }
    }
}

DVO_HD int floor_log2(int n) {
    int r = 0;
    while (n > 1) { n >>= 1; ++r; }
    return r;
}

// introselect loop with the depth budget passed in (so a parallel front end can hand over mid-way)
template <class Acc>
DVO_HDN void nth_element_replay_depth(Acc& a, int first, int nth, int last, int depth) {
    while (last - first > 3) {
        if (depth == 0) {
            heap_select(a, first, nth + 1, last);
            sel_swap(a, first, nth);
            return;
        }
        --depth;
        int mid = first + (last - first) / 2;
        move_median_to_first(a, first, first + 1, mid, last - 1);
        int cut = unguarded_partition(a, first + 1, last, first);
        if (cut <= nth) first = cut;
        else last = cut;
    }
    insertion_sort(a, first, last);
}

template <class Acc>
DVO_HDN void nth_element_replay(Acc& a, int first, int nth, int last) {
    if (first == last || nth == last) return;
    nth_element_replay_depth(a, first, nth, last, floor_log2(last - first) * 2);
}

// std::partition(first, last, [b](x){ return x >= b; }) for bidirectional iterators; returns the partition point.
template <class Acc>
DVO_HDN int partition_ge(Acc& a, int first, int last, typename Acc::Item boundary) {
    while (true) {
        first = a.scan_up_ge(first, last, boundary);     // advance while a[first] >= boundary, stop at last
        if (first == last) return first;
        --last;
        last = a.scan_down_lt(first, last, boundary);    // retreat while a[last] < boundary, stop at first
        if (first == last) return first;
        sel_swap(a, first, last);
        ++first;
    }
}

// Whole retainBest: returns the new element count; elements [0, count') are in cv2's order.
template <class Acc>
DVO_HDN int retain_best_replay(Acc& a, int count, int n_points) {
    if (!(n_points >= 0 && count > n_points)) return count;
    if (n_points == 0) return 0;
    nth_element_replay(a, 0, n_points - 1, count);
    typename Acc::Item boundary = a.get(n_points - 1);
    return partition_ge(a, n_points, count, boundary);
}

// ---------------------------------------------------------------------------------------------------------------------
// Data-parallel form of the same two algorithms.
//
// Both partition loops (Hoare's unguarded partition inside introselect, and the bidirectional std::partition) do the
// same thing: walk a left cursor to the next "left stopper", a right cursor to the next "right stopper", swap, repeat
// while left < right.  A swapped position is never looked at again in that pass, so the swaps are exactly
//     (L[k], R[k])  for k = 0 .. K-1,   K = #{k : L[k] < R[k]},
// L = left stoppers in ascending index order, R = right stoppers in descending index order, both taken on the array as
// it was BEFORE the pass.  Hoare:  left stopper = !(x > pivot), right stopper = !(pivot > x), cut = min(L[K], R[K-1]).
// std::partition(pred = x >= b):  left stopper = !pred, right stopper = pred, result = first + #right stoppers.
// That turns each pass into flags + prefix sums + K independent swaps, which a whole CTA does at once
// (orb_kernels.cu: BlockAcc); the introselect control flow around it stays the scalar one below.  tests/hostsim checks
// this formulation, with a plain-loop accessor, against the real std::nth_element / std::partition.
//
// Additional accessor contract for the paired form:
//   void median_to_first(int result, int ia, int ib, int ic)   -- move_median_to_first, then make writes visible
//   int  pair_swap_hoare(int first, int last, Item pivot)      -- returns the cut
//   int  pair_swap_ge(int first, int last, Item boundary)      -- returns the partition point
//   void sequential_tail(int first, int nth, int last, int depth) -- nth_element_replay_depth on one lane/warp + barrier
template <class Acc>
DVO_HDN void nth_element_paired(Acc& a, int first, int nth, int last, int seq_tail) {
    if (first == last || nth == last) return;
    int depth = floor_log2(last - first) * 2;
    while (last - first > 3) {
        if (depth == 0 || last - first <= seq_tail) break;
        --depth;
        int mid = first + (last - first) / 2;
        a.median_to_first(first, first + 1, mid, last - 1);
        int cut = a.pair_swap_hoare(first + 1, last, a.get(first));
        if (cut <= nth) first = cut;
        else last = cut;
    }
    a.sequential_tail(first, nth, last, depth);
}

template <class Acc>
DVO_HDN int retain_best_paired(Acc& a, int count, int n_points, int seq_tail) {
    if (!(n_points >= 0 && count > n_points)) return count;
    if (n_points == 0) return 0;
    nth_element_paired(a, 0, n_points - 1, count, seq_tail);
    typename Acc::Item boundary = a.get(n_points - 1);
    return a.pair_swap_ge(n_points, count, boundary);
}

}  // namespace dvo
