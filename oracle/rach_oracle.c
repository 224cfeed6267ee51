/*
 * rach_oracle.c -- TEST INFRASTRUCTURE (oracle).  Not part of the product path: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 *
 * A CPU restatement, in plain C, of the per-millisecond UE state machine of
 *   RandomAccessWithNOMA.c       (variant W;  loop :267-335, functions :374-415 :475-562
 *                                 :607-728 :844-847)
 *   RandomAccessSimulatorBeta.c  (variant B;  loop :111-183, functions :220-429)
 * driven by the Philox draw tape of include/rach_tape.h instead of libc rand().
 *
 * It keeps the reference's control flow (one pass over the UEs in index order per ms,
 * selectPreamble -> preambleCollision -> requestResourceAllocation -> timerIncrease) but
 * replaces the O(nUE) scan inside preambleCollision (W:613-621) by a per-preamble count
 * of the currently visible UEs (active==1 && txTime==time) and a per-preamble "collided
 * this ms" stamp that applies the reference's `txTime++` of the other group members
 * (W:653-661) lazily when their turn comes.  This makes one replication O(activeUE * ms)
 * instead of O(scans * nUE) so that 100k-UE runs take seconds, and is a formulation
 * different from the event-driven one the CUDA engine uses.
 *
 * PARITY PIN: this file is validated field by field (all 15 saveResult fields + failCount
 * of every UE, and all counters) against the reference sources themselves compiled in
 * tape mode (oracle/_ref/libref_w.so / libref_b.so, see build_ref.sh) by
 * tests/test_oracle_vs_reference.py, and against the committed fixtures under
 * tests/golden/ that were generated from those reference builds (tests/golden/make_golden.py).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>

#include "rach_tape.h"
#include "ref_api.h"

#define betaF 0.0165

typedef struct {
    int timer, active, txTime, preamble, rarWindow, maxRarCounter, preambleTxCounter;
    int msg2Flag, connectionRequest, msg4Flag, preambleChange, nowBackoff;
    int firstTxTime, secondTxTime, failCount, sector;
    int drawMs, drawK;
    float angle, x, y, distance, channelGain;
} oue;

typedef struct {
    const ref_config* cfg;
    ref_result* res;
    oue* ue;
} octx;

/* W:844-847 */
static float o_beta_dist(float a, float b, float x) {
    float betaValue = (1 / betaF) * (pow(x, (a - 1))) * (pow((1 - x), (b - 1)));
    return betaValue;
}

static int o_rand(octx* c, int i, int ms) {
    oue* u = c->ue + i;
    if (u->drawMs != ms) { u->drawMs = ms; u->drawK = 0; }
    int k = u->drawK++;
    c->res->draws++;
    if (k + 1 > c->res->maxDrawsPerUeMs) c->res->maxDrawsPerUeMs = k + 1;
    return rach_tape_rand31(c->cfg->seed, (uint32_t)c->cfg->rep, (uint32_t)i, (uint32_t)ms,
                            (uint32_t)k, RACH_TAPE_TAG_UE);
}

/* slot alignment, W:518-527 (same text at W:544-553, W:688-697) */
static int o_align(int subTime, int accessTime) {
    if (subTime % accessTime == 0) return subTime + 1;
    if (subTime % accessTime == 1) return subTime;
    return subTime + (accessTime - (subTime % accessTime) + 1);
}

/* W:383-415 (B:137-145 without the draws) */
static void o_activate(octx* c, int i, int time) {
    oue* u = c->ue + i;
    u->active = 1; u->txTime = time + 1; u->timer = 0; u->msg2Flag = 0;
    u->firstTxTime = time + 1;
    if (!c->cfg->geometry) return;
    float cellRadius = c->cfg->cellRadius;
    float bandwidth = 5;                              /* W:64 */
    float pi = 3.14;
    float theta = (float)o_rand(c, i, time) / (float)(2147483647) * 2 * pi;
    float r = cellRadius * sqrt((float)o_rand(c, i, time) / (float)2147483647);
    u->angle = theta;
    if (u->angle >= 0 && u->angle < ((1. / 3.) * pi)) u->sector = 0;
    else if (u->angle >= ((1. / 3.) * pi) && u->angle < ((2. / 3.) * pi)) u->sector = 1;
    else if (u->angle >= ((2. / 3.) * pi) && u->angle < 3.14) u->sector = 2;
    else if (u->angle >= pi && u->angle < ((4. / 3.) * pi)) u->sector = 3;
    else if (u->angle >= ((4. / 3.) * pi) && u->angle < ((5. / 3.) * pi)) u->sector = 4;
    else u->sector = 5;
    u->x = r * cos(theta);
    u->y = r * sin(theta);
    u->distance = r;
    u->channelGain = 20 * log10(4. * pi * r / (bandwidth / 1000));
}

int oracle_run(const ref_config* cfg, ref_result* res, int* perUE, float* geom) {
    const int nUE = cfg->nUE, P = cfg->nPreamble, BI = cfg->backoffIndicator;
    const int G = cfg->nGrantUL, Wn = cfg->maxRarWindow, M = cfg->maxMsg2TxCount;
    const int A = cfg->accessTime;
    if (nUE < 1 || P < 1 || BI < 1 || A < 1) return -1;
    memset(res, 0, sizeof(*res));
    struct timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);

    octx c; c.cfg = cfg; c.res = res;
    c.ue = (oue*)calloc((size_t)nUE, sizeof(oue));
    int* cnt = (int*)calloc((size_t)P, sizeof(int));        /* visible UEs per preamble   */
    int* coll = (int*)calloc((size_t)P, sizeof(int));       /* ms+1 of last collided scan */
    int* late = (int*)malloc(sizeof(int) * (size_t)nUE);    /* Msg3 restarts landing on `time` */
    if (!c.ue || !cnt || !coll || !late) return -2;
    for (int i = 0; i < nUE; ++i) {                         /* initialUE, W:374-381 */
        oue* u = c.ue + i;
        u->timer = -1; u->active = -1; u->txTime = -1; u->preamble = -1;
        u->drawMs = -1; u->sector = -1;
    }

    int maxTime, nAccessUE = 0;
    if (cfg->distribution == 1) {                           /* W:241-251 */
        maxTime = 60000;
        nAccessUE = ceil((float)nUE * (float)A * 1.0 / (float)maxTime);
        if (nAccessUE <= 0) nAccessUE = 1;
    } else {
        maxTime = 10000;                                    /* W:254 */
    }
    res->nAccessUE = nAccessUE;

    int activeCheck = 0, arrived = 0, grantCheck = 0, nSuccessUE = 0, time;
    long long continueFailed = 0, collisionPreambles = 0, totalPreambleTxop = 0;
    long long collisionScans = 0, totalScans = 0;

    for (time = 0; time < maxTime; time++) {
        if (cfg->stopMs > 0 && time >= cfg->stopMs) break;
        res->lastMs = time;
        if (time % 5 == 0) grantCheck = 0;                  /* W:268-269: literal 5 */
        if (activeCheck >= nUE) activeCheck = nUE;
        if (time % A == 0 && activeCheck != nUE) {          /* W:280-299 */
            if (cfg->distribution == 1) {
                activeCheck += nAccessUE;
            } else {
                float betaDist = o_beta_dist(3, 4, (float)time / (float)maxTime);
                int accessUEs = (int)ceil((float)nUE * betaDist / ((float)maxTime / (float)A));
                activeCheck += accessUEs;
            }
            if (activeCheck >= nUE) activeCheck = nUE;
            for (; arrived < activeCheck; ++arrived) o_activate(&c, arrived, time);
        }

        /* visible set at the start of the ms */
        memset(cnt, 0, sizeof(int) * (size_t)P);
        for (int i = 0; i < activeCheck; ++i) {
            oue* u = c.ue + i;
            if (u->active == 1 && u->txTime == time && u->msg4Flag == 0) cnt[u->preamble]++;
        }
        int nLate = 0;

        for (int i = 0; i < activeCheck; ++i) {
            oue* u = c.ue + i;
            if (u->msg4Flag != 0) continue;                 /* W:305 */
            if (u->active == 1 && u->msg2Flag == 0) {
                /* a lower-index member of my group collided earlier in this ms: W:658 */
                if (u->txTime == time && u->preamble >= 0 && coll[u->preamble] == time + 1)
                    u->txTime++;
                const int wasVisible = (u->txTime == time);
                const int oldPreamble = u->preamble;
                int moved = 0;
                /* ---- selectPreamble, W:475-562 ---- */
                if (u->preamble == -1) {
                    u->preamble = o_rand(&c, i, time) % P;
                    u->rarWindow = 0; u->maxRarCounter = 0; u->preambleChange = 1;
                    u->preambleTxCounter = 1; u->nowBackoff = 0; u->failCount = 0;
                } else if (u->nowBackoff <= 0) {
                    u->rarWindow++;
                    if (u->rarWindow >= Wn) {
                        moved = 1;
                        if (u->maxRarCounter >= M) {        /* limit branch, W:498-531 */
                            continueFailed++;
                            u->preamble = o_rand(&c, i, time) % P;
                            u->rarWindow = 0; u->maxRarCounter = 0; u->preambleChange = 1;
                            u->preambleTxCounter = 1; u->nowBackoff = 0; u->timer = 0;
                            u->firstTxTime = time + 1; u->failCount += 1;
                            int tmp = o_rand(&c, i, time) % BI;
                            u->txTime = o_align(u->txTime + tmp, A);   /* W:516: CURRENT txTime */
                            u->nowBackoff = u->txTime - time;
                        } else {                            /* retry branch, W:532-558 */
                            u->rarWindow = 0; u->maxRarCounter++; u->preambleTxCounter++;
                            int tmp = o_rand(&c, i, time) % BI;
                            u->txTime = o_align(time + tmp, A);
                            u->nowBackoff = u->txTime - time;
                            u->secondTxTime = u->txTime;
                        }
                    }
                }
                if (moved) {
                    if (wasVisible) cnt[oldPreamble]--;
                    if (u->txTime == time) cnt[u->preamble]++;
                }
                /* ---- preambleCollision, W:310-314, 607-665 ---- */
                if (u->txTime == time) {
                    int check = cnt[u->preamble];
                    if (check < 1) { res->aborted = 99; check = 1; }   /* invariant I4 */
                    totalScans++;
                    if (check == 1) {
                        totalPreambleTxop++;
                        grantCheck++;
                        if (grantCheck < G) {
                            u->active = 2; u->txTime = time + 11;
                            u->connectionRequest = 0; u->msg2Flag = 1;
                        } else {
                            u->txTime++;
                        }
                    } else {
                        collisionPreambles += check; totalPreambleTxop += check;
                        collisionScans++;
                        u->txTime++;
                        coll[u->preamble] = time + 1;       /* unprocessed members: lazily */
                        for (int k = 0; k < nLate; ++k) {   /* already processed members   */
                            oue* v = c.ue + late[k];
                            if (v->active == 1 && v->txTime == time && v->preamble == u->preamble) {
                                v->txTime++; res->lateAbsorbed++;
                            }
                        }
                    }
                    cnt[u->preamble] = 0;
                }
            }
            /* ---- requestResourceAllocation, W:318-321, 667-710 ---- */
            if (u->active == 2 && u->txTime == time) {
                u->connectionRequest++;
                if (u->connectionRequest < 48) {
                    int r = o_rand(&c, i, time);
                    if (rach_msg3_success(r)) {
                        u->msg4Flag = 1; u->timer = u->timer + 6; u->active = 0;
                        nSuccessUE++;
                    } else {
                        u->connectionRequest = 48; u->txTime += 48;
                    }
                } else {
                    continueFailed++;
                    int tmp = o_rand(&c, i, time) % BI;
                    u->txTime = o_align(u->txTime + tmp, 5);          /* W:687: literal 5 */
                    u->active = 1;
                    u->nowBackoff = u->txTime - time;
                    u->preamble = o_rand(&c, i, time) % P;
                    u->timer = 0; u->msg2Flag = 0; u->rarWindow = 0; u->maxRarCounter = 0;
                    u->connectionRequest = 0; u->failCount += 1;
                    if (u->txTime == time) { cnt[u->preamble]++; late[nLate++] = i; res->lateRestarts++; }
                }
            }
            /* ---- timerIncrease, W:324-325, 712-718 ---- */
            if (u->active > 0) {
                u->timer++;
                if (u->nowBackoff > 0) u->nowBackoff--;
            }
        }
        if (nSuccessUE == nUE) break;                       /* W:330-334 */
    }

    long long txSum = 0, delaySum = 0, failSum = 0;
    for (int i = 0; i < nUE; ++i) {
        oue* u = c.ue + i;
        if (u->msg4Flag == 1) { txSum += u->preambleTxCounter; delaySum += u->timer; failSum += u->failCount; }
        if (perUE) {
            int* o = perUE + (size_t)i * 16;
            o[0] = u->timer; o[1] = u->active; o[2] = u->txTime; o[3] = u->firstTxTime;
            o[4] = u->secondTxTime; o[5] = u->nowBackoff; o[6] = u->preamble;
            o[7] = u->preambleChange; o[8] = u->rarWindow; o[9] = u->maxRarCounter;
            o[10] = u->preambleTxCounter; o[11] = u->msg2Flag; o[12] = u->connectionRequest;
            o[13] = u->msg4Flag; o[14] = u->failCount;
            o[15] = (cfg->geometry && u->active != -1) ? u->sector : -1;
        }
        if (geom) {
            float* o = geom + (size_t)i * 6;
            o[0] = u->angle; o[1] = u->x; o[2] = u->y; o[3] = u->distance; o[4] = u->channelGain;
            o[5] = (float)(u->active == -1 ? 0 : (cfg->geometry ? u->sector : 0));
        }
    }
    res->simTimeMs = time; res->nSuccess = nSuccessUE;
    res->preambleTxSum = txSum; res->delaySum = delaySum;
    res->failCountSum = failSum;
    res->continueFailed = continueFailed;
    res->collisionPreambles = collisionPreambles; res->totalPreambleTxop = totalPreambleTxop;
    res->collisionScans = collisionScans; res->totalScans = totalScans;
    res->failCountsPrinted = (int)failSum;
    res->captured = 1;
    clock_gettime(CLOCK_MONOTONIC, &t1);
    res->seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    free(c.ue); free(cnt); free(coll); free(late);
    return 0;
}

/* the Beta arrival counts of W:285-287 alone (C semantics), for the host-schedule test */
int oracle_beta_arrivals(int nUE, int accessTime, int* perMs /* [10000] */) {
    const int maxTime = 10000;
    int activeCheck = 0, allAt = -1;
    for (int time = 0; time < maxTime; ++time) {
        perMs[time] = 0;
        if (activeCheck >= nUE) activeCheck = nUE;
        if (time % accessTime == 0 && activeCheck != nUE) {
            const int before = activeCheck;
            float betaDist = o_beta_dist(3, 4, (float)time / (float)maxTime);
            int accessUEs = (int)ceil((float)nUE * betaDist / ((float)maxTime / (float)accessTime));
            activeCheck += accessUEs;
            if (activeCheck >= nUE) activeCheck = nUE;
            perMs[time] = activeCheck - before;
            if (activeCheck == nUE && allAt < 0) allAt = time;
        }
    }
    return allAt;
}

int oracle_sizeof_config(void) { return (int)sizeof(ref_config); }
int oracle_sizeof_result(void) { return (int)sizeof(ref_result); }
