/*
 * rach_core_u0.cuh -- variant U0: the oldest simulator, RandomAccessSimulator.c (main loop
 * U0:75-126, selectPreamble U0:157-197, preambleCollision U0:200-233,
 * requestResourceAllocation U0:235-257, timerIncrease U0:259-266).
 *
 * U0 has index-order effects that differ from W (the group update of U0:228-229 makes the
 * members above the scanner back off in the same ms and the scanner one ms later; UEs in the
 * Msg3 phase scan too, U0:107; dropped UEs keep colliding, U0:207 has no raFailed test), a
 * 60 s Uniform horizon and a live set of a few dozen UEs.  It is run here the simple exact way:
 * ONE THREAD PER REPLICATION walks the ms loop over a compact, index-ordered list of the live
 * UEs (arrived, not finished; dropped UEs stay as "phantoms") and performs the reference's own
 * steps on it -- the O(nUE) loops of the reference shrink to O(live).  Replications are
 * independent, so a launch runs thousands of them side by side, each in a warp of its own (the walks
 * diverge completely), the head of the live list in shared memory.  A lane-parallel formulation like W's
 * is the next step for this variant.
 */
#ifndef RACH_CORE_U0_CUH
#define RACH_CORE_U0_CUH

#include "rach_core.cuh"

#ifdef __CUDACC__
#define RU_ALIGN __align__(16)
#else
#define RU_ALIGN alignas(16)
#endif
struct RU_ALIGN RuUE {   /* 64 bytes; the first 16 are all a collision scan reads (U0:207-209) */
    int idx, active, txTime, preamble;      /* live-list copy: active == -2 = gone (finished / moved to ph[]) */
    int timer, rarWindow, maxRarCounter, preambleTxCounter;
    int msg2Flag, connectionRequest, msg4Flag, preambleChange;
    int raFailed, nowBackoff, pad0, pad1;   /* pad0: next node of a phantom chain */
};

struct RuStats { int simTime, nSuccess; long long txSum, delaySum, collisionPreambles, totalPreambleTxop, dropped; int overflow; };

/* dump row: timer active txTime preamble preambleChange rarWindow maxRarCounter preambleTxCounter
 * msg2Flag connectionRequest msg4Flag raFailed nowBackoff 0 0 0 (saveResult order, U0:339-341) */
RA_HD void ru_dump_row(int* o, const RuUE& u) {
    o[0] = u.timer; o[1] = u.active; o[2] = u.txTime; o[3] = u.preamble; o[4] = u.preambleChange;
    o[5] = u.rarWindow; o[6] = u.maxRarCounter; o[7] = u.preambleTxCounter; o[8] = u.msg2Flag;
    o[9] = u.connectionRequest; o[10] = u.msg4Flag; o[11] = u.raFailed; o[12] = u.nowBackoff;
    o[13] = 0; o[14] = 0; o[15] = 0;
}


/* Dropped UEs ("phantoms", raFailed == -1) are never processed again (U0:99) but still match the scan
 * predicate of U0:207 in the one ms where txTime+2 == time; a group update then moves them to time+3
 * (U0:228-229) so they can match again 5 ms later.  They are kept out of the live list in `ph[]`,
 * chained per ms of their next possible match (pad0 = next node, -1 ends the chain). */
/* The first `winCap` entries of the live list live in `win` (shared memory on the device: the live set is a few dozen
 * UEs, and a thread that walks it every ms out of L2 spends its time waiting), the rest in `live` (global). */
#define RU_L(x) (*((x) < winCap ? win + (x) : live + (x)))
template <bool DUMP>
RA_HD void ru_run_replication(const RaJob& job, RuUE* live, RuUE* win, int winCap, RuUE* ph, int* phHead, int cap, RuStats* out) {
    const RaPointDev& pt = *job.pt;
    const int nUE = pt.nUE, P = pt.P, BI = pt.BI, maxTime = pt.maxTime, accessTime = 5;   /* U0:57,59 */
    const int nAccessUE = pt.G;                       /* host: ceil(n*5/60000), at least 1 (U0:60-64) */
    const int ringMask = pt.R - 1;
    RuStats st; st.simTime = maxTime; st.nSuccess = 0; st.txSum = st.delaySum = 0;
    st.collisionPreambles = st.totalPreambleTxop = st.dropped = 0; st.overflow = 0;
    if (DUMP) for (int i = 0; i < nUE; ++i) {         /* calloc + initialUE, U0:46,149-155 */
        int* o = job.dump + (size_t)i * RA_DUMP_W;
        for (int k = 0; k < RA_DUMP_W; ++k) o[k] = 0;
        o[0] = -1; o[1] = -1; o[2] = -1; o[3] = -1;
    }
    for (int i = 0; i <= ringMask; ++i) phHead[i] = -1;
    int nLive = 0, nGone = 0, nPh = 0, activeCheck = 0, arrived = 0, time;
    for (time = 0; time < maxTime; time++) {
        if (time % accessTime == 1) {                 /* U0:77-94 (bound fixed: i < nUE) */
            if (activeCheck >= nUE) activeCheck = nUE; else activeCheck += nAccessUE;
            int upto = activeCheck + 1 < nUE ? activeCheck + 1 : nUE;
            for (; arrived < upto; ++arrived) {
                if (nLive >= cap) { st.overflow = 1; break; }
                RuUE u; u.idx = arrived; u.timer = 0; u.active = 1; u.txTime = time + 1; u.preamble = -1;
                u.rarWindow = 0; u.maxRarCounter = 0; u.preambleTxCounter = 0; u.msg2Flag = 0;
                u.connectionRequest = 0; u.msg4Flag = 0; u.preambleChange = 0; u.raFailed = 0; u.nowBackoff = 0;
                u.pad0 = u.pad1 = 0;
                RU_L(nLive) = u; ++nLive;
            }
        }
        const int slot = time & ringMask;
        for (int a = 0; a < nLive; ++a) {
            RuUE u = RU_L(a);
            if (u.active == -2) continue;                                     /* finished or moved to ph[] */
            unsigned k = 0;
            if (u.active == 1 && u.msg2Flag == 0) {                           /* selectPreamble U0:157-197 */
                if (u.preamble == -1) {
                    u.preamble = (int)ra_mod((unsigned)rach_tape_rand31(pt.seed, job.rep, (unsigned)u.idx, (unsigned)time, k++, RACH_TAPE_TAG_UE), (unsigned)P, pt.magicP);
                    u.rarWindow = 0; u.maxRarCounter = 0; u.preambleTxCounter = 0; u.preambleChange = 1;
                } else if (u.nowBackoff == 0) {
                    u.rarWindow++;
                    if (u.rarWindow >= 5) {
                        int tmp = (int)ra_mod((unsigned)rach_tape_rand31(pt.seed, job.rep, (unsigned)u.idx, (unsigned)time, k++, RACH_TAPE_TAG_UE), (unsigned)BI, pt.magicBI) + 2;
                        u.txTime = time + tmp; u.nowBackoff = tmp; u.rarWindow = 0; u.maxRarCounter++;
                        if (u.maxRarCounter >= 10) {
                            u.raFailed = -1; st.dropped++;
                            u.preamble = (int)ra_mod((unsigned)rach_tape_rand31(pt.seed, job.rep, (unsigned)u.idx, (unsigned)time, k++, RACH_TAPE_TAG_UE), (unsigned)P, pt.magicP);
                            u.maxRarCounter = 0; u.preambleChange++;
                        }
                    }
                }
            }
            if (u.txTime + 2 == time && u.txTime != -1) {                     /* preambleCollision U0:107-110, 200-233 */
                RU_L(a) = u;                                                  /* the scan below reads the list */
                int check = 0;
                for (int b = 0; b < nLive; ++b)
                    { const RuUE& o = RU_L(b); if (o.active == 1 && o.txTime + 2 == time && o.preamble == u.preamble) check++; }
                for (int n = phHead[slot]; n >= 0; n = ph[n].pad0)
                    if (ph[n].txTime + 2 == time && ph[n].preamble == u.preamble) check++;
                if (check == 1) {
                    st.totalPreambleTxop++;
                    u.preambleTxCounter++; u.active = 2; u.txTime = time + 2; u.connectionRequest = 0; u.msg2Flag = 1;
                } else {
                    st.collisionPreambles += check;
                    for (int b = 0; b < nLive; ++b)
                    {
                        RuUE& o = RU_L(b);
                        if (o.active == 1 && o.txTime + 2 == time && o.preamble == u.preamble) { o.rarWindow = 5; o.txTime = time + 3; }
                    }
                    /* phantoms of this class: same update, then they can match again at time+5 */
                    int prev = -1;
                    for (int n = phHead[slot]; n >= 0;) {
                        const int next = ph[n].pad0;
                        if (ph[n].txTime + 2 == time && ph[n].preamble == u.preamble) {
                            ph[n].rarWindow = 5; ph[n].txTime = time + 3;
                            if (prev < 0) phHead[slot] = next; else ph[prev].pad0 = next;
                            ph[n].pad0 = phHead[(time + 5) & ringMask]; phHead[(time + 5) & ringMask] = n;
                        } else prev = n;
                        n = next;
                    }
                    u = RU_L(a);                                              /* the scanner may be a member */
                }
            }
            if (u.active == 2 && u.txTime + 2 == time) {                      /* requestResourceAllocation U0:113-115, 235-257 */
                u.connectionRequest++;
                if (u.connectionRequest < 48) {
                    int r = rach_tape_rand31(pt.seed, job.rep, (unsigned)u.idx, (unsigned)time, k++, RACH_TAPE_TAG_UE);
                    if (rach_msg3_success(r)) { u.msg4Flag = 1; u.active = 0; st.nSuccess++; }
                    else u.txTime = time + 1;
                } else {
                    u.active = 1;
                    u.txTime = time + (int)ra_mod((unsigned)rach_tape_rand31(pt.seed, job.rep, (unsigned)u.idx, (unsigned)time, k++, RACH_TAPE_TAG_UE), (unsigned)BI, pt.magicBI) + 2;
                    u.preamble = (int)ra_mod((unsigned)rach_tape_rand31(pt.seed, job.rep, (unsigned)u.idx, (unsigned)time, k++, RACH_TAPE_TAG_UE), (unsigned)P, pt.magicP);
                    u.msg2Flag = 0; u.rarWindow = 0; u.maxRarCounter = 0; u.preambleTxCounter++;
                }
            }
            if (u.active > 0 && u.msg4Flag == 0) {                            /* U0:118-119, 259-266 */
                u.timer++;
                if (u.nowBackoff != 0) u.nowBackoff--;
            }
            if (u.msg4Flag == 1) {
                st.txSum += u.preambleTxCounter; st.delaySum += u.timer; nGone++;
                if (DUMP) ru_dump_row(job.dump + (size_t)u.idx * RA_DUMP_W, u);
                u.active = -2;
            } else if (u.raFailed == -1) {                                    /* frozen from now on: to the phantom calendar */
                if (nPh >= cap) { st.overflow = 1; }
                else {
                    RuUE n = u; const int s2 = (u.txTime + 2) & ringMask;
                    n.pad0 = phHead[s2]; ph[nPh] = n; phHead[s2] = nPh; nPh++;
                }
                nGone++; u.active = -2;
            }
            RU_L(a) = u;
        }
        phHead[slot] = -1;                                                    /* whoever was not hit never matches again */
        if (st.nSuccess == nUE) break;                                        /* U0:122-125 */
        if (nGone > 16 && nGone * 4 > nLive) {                                /* compact, keeping index order */
            int w = 0;
            for (int a = 0; a < nLive; ++a) if (RU_L(a).active != -2) { if (w != a) RU_L(w) = RU_L(a); ++w; }
            nLive = w; nGone = 0;
        }
    }
    st.simTime = time;
    if (DUMP) {
        for (int a = 0; a < nLive; ++a) if (RU_L(a).active != -2) ru_dump_row(job.dump + (size_t)RU_L(a).idx * RA_DUMP_W, RU_L(a));
        for (int n = 0; n < nPh; ++n) ru_dump_row(job.dump + (size_t)ph[n].idx * RA_DUMP_W, ph[n]);
    }
    *out = st;
}

#undef RU_L

#endif /* RACH_CORE_U0_CUH */
