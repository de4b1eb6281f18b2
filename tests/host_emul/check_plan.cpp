// TEST-ONLY: exercises the shipped mode-B schedule planner (miniraytracer_b200/csrc/schedule.h) on the CPU.
// stdin: lines "n_pixels ns resident_warps has_trees coop chunk_pixels chunk_paths tail_tasks max_items"; for each line the tickets
// are walked with the kernel's own task -> pixels mapping and must tile [0, n_pixels) in order, without gaps or overlaps, every
// chunk within the staging capacity.  Prints one JSON line per case; exit code 1 on the first violation.
#include <cstdio>
#include <cstring>

#include "schedule.h"

int main() {
    unsigned long long n_pixels, ns, warps, trees, coop, cpx, cpaths, tail, max_items;
    int line = 0;
    while (scanf("%llu %llu %llu %llu %llu %llu %llu %llu %llu", &n_pixels, &ns, &warps, &trees, &coop, &cpx, &cpaths, &tail, &max_items) == 9) {
        line++;
        MrtTuning tn;
        memset(&tn, 0, sizeof(tn));
        tn.chunk_pixels = (uint32_t) cpx; tn.chunk_paths = (uint32_t) cpaths; tn.tail_tasks = (uint32_t) tail;
        const mrt::BinnedPlan p = mrt::plan_binned_schedule((uint32_t) n_pixels, (uint32_t) ns, (uint32_t) warps, trees != 0, coop != 0, tn, (uint32_t) max_items);
        unsigned long long next = 0, biggest = 0, smallest_last = 0;
        for (uint32_t t = 0; t < p.n_tasks; t++) {
            uint32_t pix0, kp;
            mrt::plan_task_pixels(p, t, &pix0, &kp);
            if (pix0 != next || kp == 0 || (unsigned long long) kp * ns > max_items || pix0 + kp > n_pixels) {
                printf("{\"line\": %d, \"error\": \"task %u covers [%u, %u) but the next free pixel is %llu\"}\n", line, t, pix0, pix0 + kp, next);
                return 1;
            }
            next += kp;
            if (kp > biggest) biggest = kp;
            smallest_last = kp;
        }
        if (next != n_pixels) { printf("{\"line\": %d, \"error\": \"tasks cover %llu of %llu pixels\"}\n", line, next, n_pixels); return 1; }
        if (!(p.task0[0] == 0 && p.task0[1] <= p.task0[2] && p.task0[2] <= p.n_tasks && p.pix0[1] <= p.pix0[2] && p.pix0[2] <= p.pix0[3] && p.pix0[3] == n_pixels)) {
            printf("{\"line\": %d, \"error\": \"runs out of order\"}\n", line);
            return 1;
        }
        printf("{\"line\": %d, \"K\": %u, \"n_tasks\": %u, \"k\": [%u, %u, %u], \"pix0\": [%u, %u, %u, %u], \"biggest\": %llu, \"last\": %llu}\n", line, p.K, p.n_tasks,
               p.k[0], p.k[1], p.k[2], p.pix0[0], p.pix0[1], p.pix0[2], p.pix0[3], biggest, smallest_last);
    }
    return 0;
}
