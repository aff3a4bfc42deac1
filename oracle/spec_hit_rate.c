/*
 * spec_hit_rate.c -- TEST INFRASTRUCTURE / design experiment (not part of the product, not part of liborc.so).
 *
 * Question (DESIGN.md "Known gaps", speculative selection): with wave-1 semantics (mcts.py:348-392 with batch_size 1)
 * simulation i+1 can only start once simulation i's rollout value v in {-1, 0, +1} is known.  A kernel could decide
 * the NEXT leaf under each hypothetical v while the rollout runs and start ONE speculative rollout early.  How often
 * would that speculative rollout be the right one?
 *
 * This program replays the oracle's wave-1 search (rollout evaluator) and, before every real backup, clones the tree,
 * backs up each hypothetical v, re-runs the traverse and records the leaf it reaches.
 *
 *   gcc -O2 -std=c11 -ffp-contract=off -o /tmp/spec_hit_rate oracle/spec_hit_rate.c -lm && /tmp/spec_hit_rate
 */
#include "rvs_oracle.c"

#include <stdio.h>

static int traverse_clone(const otree *src, const int *path, int plen, float hv, const orc_board *root, int rules,
                          onode *scratch) {
    otree t = *src;
    memcpy(scratch, src->nodes, sizeof(onode) * (size_t)src->n);
    t.nodes = scratch;
    backprop(&t, path, plen, hv);
    orc_board b = *root;
    int node = 0;
    while (t.nodes[node].nchild > 0 && !t.nodes[node].terminal) {
        t.nodes[node].VL++;
        float best = -INFINITY;
        int next = -1;
        int fc = t.nodes[node].first_child, nc = t.nodes[node].nchild, pn = t.nodes[node].N;
        for (int c = fc; c < fc + nc; c++) {
            float s = ucb(&t, c, pn);
            if (s > best) { best = s; next = c; }
        }
        if (next < 0) return -1;
        orc_apply(&b, t.nodes[next].move, rules);
        node = next;
    }
    return node;
}

int main(void) {
    const int rules = 0, S = 100;
    long sims = 0, same_all = 0, same_pm = 0, hit_plus = 0, hit_minus = 0, hit_prev = 0, hit_same_as_cur = 0;
    long rollouts = 0, draws = 0;
    long by_phase[7][3] = {{0}};
    for (int g = 0; g < 600; g++) {
        orc_board root;
        orc_board_init(&root);
        /* a root at a random phase: g % 56 uniform-random plies from the start */
        uint32_t rs = (uint32_t)orc_stream_seed(99, (uint64_t)g, 1);
        for (int p = 0; p < g % 56 && !root.over; p++) {
            uint64_t lm = orc_board_legal(&root, rules);
            if (!lm) break;
            orc_apply(&root, nth_set_bit(lm, (int)(((uint64_t)roll_next(&rs) * (uint64_t)popc(lm)) >> 32)), rules);
        }
        if (root.over) continue;
        const int phase = popc(root.black | root.white) / 10;
        otree t;
        t.cap = 2 + S * 34;
        t.nodes = (onode *)malloc(sizeof(onode) * (size_t)t.cap);
        onode *scratch = (onode *)malloc(sizeof(onode) * (size_t)t.cap);
        t.n = 0;
        t.c_puct = 1.0f;
        tree_new(&t, 1.0f, root.side, 255);
        int pred[3] = {-2, -2, -2}; /* leaves predicted by the previous simulation for v = -1, 0, +1 */
        float prev_v = 1.0f;
        for (int s = 0; s < S; s++) {
            orc_board b = root;
            int path[64], plen = 0, node = 0;
            path[plen++] = 0;
            while (t.nodes[node].nchild > 0 && !t.nodes[node].terminal) {
                t.nodes[node].VL++;
                float best = -INFINITY;
                int next = -1;
                int fc = t.nodes[node].first_child, nc = t.nodes[node].nchild, pn = t.nodes[node].N;
                for (int c = fc; c < fc + nc; c++) {
                    float sc = ucb(&t, c, pn);
                    if (sc > best) { best = sc; next = c; }
                }
                orc_apply(&b, t.nodes[next].move, rules);
                node = next;
                path[plen++] = node;
            }
            (void)pred;
            float v;
            int rolled = 0;
            if (t.nodes[node].terminal) {
                v = t.nodes[node].term_value;
            } else {
                uint64_t lm = orc_board_legal(&b, rules);
                if (lm == 0) {
                    onode *x = &t.nodes[node];
                    x->terminal = 1;
                    x->term_value = !b.over ? 0.0f : b.winner == 1 ? 1.0f : b.winner == 2 ? -1.0f : 0.0f;
                    v = x->term_value;
                } else {
                    orc_board c = b;
                    orc_random_playout(&c, orc_stream_seed(7, (uint64_t)g, (uint64_t)s), rules);
                    v = (!c.over || c.winner == 0) ? 0.0f : (c.winner == b.side ? 1.0f : -1.0f);
                    rolled = 1;
                    int fc = t.n, nc = popc(lm);
                    while (lm) {
                        int sq = __builtin_ctzll(lm);
                        lm &= lm - 1;
                        tree_new(&t, 1.0f / 65.0f, 3 - t.nodes[node].turn, sq);
                    }
                    t.nodes[node].first_child = fc;
                    t.nodes[node].nchild = nc;
                }
            }
            if (rolled && s + 1 < S) {
                /* the speculation window: this simulation's rollout.  Leaves of simulation s+1 under each v: */
                int lf[3];
                for (int h = 0; h < 3; h++) lf[h] = traverse_clone(&t, path, plen, (float)(h - 1), &root, rules, scratch);
                const int actual = lf[(int)v + 1];
                sims++;
                rollouts++;
                if (v == 0.0f) draws++;
                if (lf[0] == lf[2] && lf[1] == lf[0]) same_all++;
                if (lf[0] == lf[2]) { same_pm++; by_phase[phase][0]++; }
                by_phase[phase][1]++;
                if (lf[2] == actual) hit_plus++;
                if (lf[0] == actual) hit_minus++;
                if (lf[(int)prev_v + 1] == actual) hit_prev++;
                (void)hit_same_as_cur;
            }
            backprop(&t, path, plen, v);
            prev_v = v;
        }
        free(t.nodes);
        free(scratch);
    }
    printf("speculation windows (simulations with a rollout, not the last of a search): %ld (draw rollouts: %ld)\n", rollouts, draws);
    printf("next leaf identical for v=-1 and v=+1         : %.3f\n", (double)same_pm / sims);
    printf("next leaf identical for v=-1, 0, +1           : %.3f\n", (double)same_all / sims);
    printf("hit rate of 'assume v=+1'                     : %.3f\n", (double)hit_plus / sims);
    printf("hit rate of 'assume v=-1'                     : %.3f\n", (double)hit_minus / sims);
    printf("hit rate of 'assume the previous rollout's v' : %.3f\n", (double)hit_prev / sims);
    for (int p = 0; p < 7; p++)
        if (by_phase[p][1]) printf("  discs %2d-%2d: v-independent next leaf %.3f (%ld windows)\n", p * 10, p * 10 + 9, (double)by_phase[p][0] / by_phase[p][1], by_phase[p][1]);
    return 0;
}
