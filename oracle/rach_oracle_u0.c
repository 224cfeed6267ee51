/*
 * rach_oracle_u0.c -- TEST INFRASTRUCTURE (oracle).  CPU restatement of the oldest simulator,
 * RandomAccessSimulator.c (variant U0): main loop U0:75-126, selectPreamble U0:157-197,
 * preambleCollision U0:200-233, requestResourceAllocation U0:235-257, timerIncrease U0:259-266,
 * with the two fixes the tape-mode reference build also applies (array bound at U0:84, the
 * redeclared local at U0:321 is outside the state machine).  Draw tape: include/rach_tape.h.
 *
 * Same control flow as the reference (one pass over ALL nUE per ms in index order,
 * selectPreamble -> collision check when txTime+2 == time -> Msg3 -> timers) with the O(nUE)
 * scan of preambleCollision replaced by per-preamble counts of the UEs that match the scan
 * predicate (active==1 && txTime+2==time, U0:207 -- note: no raFailed test, so dropped UEs keep
 * colliding as "phantoms") and a lazily applied group update (rarWindow=5, txTime=time+3, U0:228-229).
 *
 * PARITY PIN: validated against RandomAccessSimulator.c itself in tape mode
 * (oracle/_ref/libref_u0.so) by tests/test_oracle_vs_reference.py.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "rach_tape.h"
#include "ref_api.h"

typedef struct {
    int timer, active, txTime, preamble, rarWindow, maxRarCounter, preambleTxCounter, msg2Flag;
    int connectionRequest, msg4Flag, preambleChange, raFailed, nowBackoff;
    int drawMs, drawK;
} uue;

typedef struct { const ref_config* cfg; ref_result* res; uue* ue; } uctx;

static int u_rand(uctx* c, int i, int ms) {
    uue* u = c->ue + i;
    if (u->drawMs != ms) { u->drawMs = ms; u->drawK = 0; }
    int k = u->drawK++;
    c->res->draws++;
    if (k + 1 > c->res->maxDrawsPerUeMs) c->res->maxDrawsPerUeMs = k + 1;
    return rach_tape_rand31(c->cfg->seed, (uint32_t)c->cfg->rep, (uint32_t)i, (uint32_t)ms, (uint32_t)k, RACH_TAPE_TAG_UE);
}

int oracle_run_u0(const ref_config* cfg, ref_result* res, int* perUE, float* geom) {
    (void)geom;
    const int nUE = cfg->nUE, P = cfg->nPreamble, BI = cfg->backoffIndicator;
    if (nUE < 1 || P < 1 || BI < 1) return -1;
    memset(res, 0, sizeof(*res));
    struct timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
    uctx c; c.cfg = cfg; c.res = res;
    c.ue = (uue*)calloc((size_t)nUE, sizeof(uue));
    int* cnt = (int*)calloc((size_t)P, sizeof(int));       /* UEs matching U0:207 per preamble */
    int* coll = (int*)calloc((size_t)P, sizeof(int));      /* ms+1 of the last group update    */
    int* phantom = (int*)malloc(sizeof(int) * (size_t)nUE);/* dropped UEs matching in this ms  */
    for (int i = 0; i < nUE; ++i) {                         /* initialUE U0:149-155 */
        uue* u = c.ue + i;
        u->timer = -1; u->active = -1; u->txTime = -1; u->preamble = -1; u->drawMs = -1;
    }
    const int maxTime = 60000, accessTime = 5;              /* U0:57,59 */
    int nAccessUE = ceil((float)nUE * (float)accessTime * 1.0 / (float)maxTime);   /* U0:60 */
    if (nAccessUE == 0) nAccessUE = 1;
    int activeCheck = 0, nSuccess = 0, time;
    int lo = 0, hi = 0;   /* UEs below lo have all finished, UEs from hi on have not arrived: both inert */
    long long collisionPreambles = 0, totalPreambleTxop = 0;

    for (time = 0; time < maxTime; time++) {
        if (cfg->stopMs > 0 && time >= cfg->stopMs) break;
        res->lastMs = time;
        if (time % accessTime == 1) {                       /* U0:77-94 */
            if (activeCheck >= nUE) activeCheck = nUE; else activeCheck += nAccessUE;
            for (int i = 0; i <= activeCheck && i < nUE; i++) {
                uue* u = c.ue + i;
                if (u->active == -1) { u->active = 1; u->txTime = time + 1; u->timer = 0; u->msg2Flag = 0; }
            }
            hi = activeCheck + 1 < nUE ? activeCheck + 1 : nUE;
        }
        while (lo < hi && c.ue[lo].msg4Flag == 1) lo++;
        memset(cnt, 0, sizeof(int) * (size_t)P);
        int nPh = 0;
        for (int i = lo; i < hi; ++i) {
            uue* u = c.ue + i;
            if (u->active == 1 && u->txTime + 2 == time && u->preamble >= 0) {
                cnt[u->preamble]++;
                if (u->raFailed == -1) phantom[nPh++] = i;
            }
        }
        for (int i = lo; i < hi; ++i) {
            uue* u = c.ue + i;
            if (!(u->msg4Flag == 0 && u->raFailed != -1)) continue;       /* U0:99 */
            /* group update by a lower-index scanner of my preamble, U0:228-229 */
            if (u->active == 1 && u->txTime + 2 == time && u->preamble >= 0 && coll[u->preamble] == time + 1) {
                u->rarWindow = 5; u->txTime = time + 3;
            }
            int wasMember = (u->active == 1 && u->txTime + 2 == time && u->preamble >= 0);
            const int oldP = u->preamble;
            if (u->active == 1 && u->msg2Flag == 0) {        /* selectPreamble U0:157-197 */
                if (u->preamble == -1) {
                    u->preamble = u_rand(&c, i, time) % P;
                    u->rarWindow = 0; u->maxRarCounter = 0; u->preambleTxCounter = 0; u->preambleChange = 1;
                } else if (u->nowBackoff == 0) {
                    u->rarWindow++;
                    if (u->rarWindow >= 5) {
                        int tmp = (u_rand(&c, i, time) % BI) + 2;
                        u->txTime = time + tmp; u->nowBackoff = tmp; u->rarWindow = 0; u->maxRarCounter++;
                        if (u->maxRarCounter >= 10) {
                            u->raFailed = -1;
                            u->preamble = u_rand(&c, i, time) % P;
                            u->maxRarCounter = 0; u->preambleChange++;
                        }
                    }
                }
            }
            if (wasMember && !(u->active == 1 && u->txTime + 2 == time && u->preamble == oldP)) cnt[oldP]--;
            /* preambleCollision U0:107-110, 200-233: any UE whose txTime+2 == time scans, whatever its phase */
            if (u->txTime + 2 == time && u->txTime != -1) {
                const int p = u->preamble;
                int check = cnt[p];
                if (check == 1) {
                    totalPreambleTxop++;
                    u->preambleTxCounter++; u->active = 2; u->txTime = time + 2; u->connectionRequest = 0; u->msg2Flag = 1;
                    /* if the scanner was itself the transmitter it has left the class; a Msg3 visitor
                       (active was 2) leaves the real transmitter in place */
                    if (wasMember && oldP == p) cnt[p] = 0;
                } else {
                    collisionPreambles += check;
                    if (check > 0) {
                        coll[p] = time + 1;
                        for (int k = 0; k < nPh; ++k) {
                            uue* v = c.ue + phantom[k];
                            if (v->active == 1 && v->txTime + 2 == time && v->preamble == p) { v->rarWindow = 5; v->txTime = time + 3; }
                        }
                        if (u->active == 1 && wasMember) { u->rarWindow = 5; u->txTime = time + 3; }
                        cnt[p] = 0;
                    }
                }
            }
            if (u->active == 2 && u->txTime + 2 == time) {   /* requestResourceAllocation U0:113-115, 235-257 */
                u->connectionRequest++;
                if (u->connectionRequest < 48) {
                    int r = u_rand(&c, i, time);
                    if (rach_msg3_success(r)) { u->msg4Flag = 1; u->active = 0; nSuccess++; }
                    else u->txTime = time + 1;
                } else {
                    u->active = 1;
                    u->txTime = time + (u_rand(&c, i, time) % BI) + 2;
                    u->preamble = u_rand(&c, i, time) % P;
                    u->msg2Flag = 0; u->rarWindow = 0; u->maxRarCounter = 0; u->preambleTxCounter++;
                }
            }
            if (u->active > 0 && u->msg4Flag == 0) {         /* U0:118-119, 259-266 */
                if (u->active != 0) u->timer++;
                if (u->nowBackoff != 0) u->nowBackoff--;
            }
        }
        if (nSuccess == nUE) break;                          /* U0:122-125 */
    }
    long long txSum = 0, delaySum = 0;
    for (int i = 0; i < nUE; ++i) {
        uue* u = c.ue + i;
        if (u->msg4Flag == 1) { txSum += u->preambleTxCounter; delaySum += u->timer; }
        if (perUE) {
            int* o = perUE + (size_t)i * 16;
            o[0] = u->timer; o[1] = u->active; o[2] = u->txTime; o[3] = u->preamble; o[4] = u->preambleChange;
            o[5] = u->rarWindow; o[6] = u->maxRarCounter; o[7] = u->preambleTxCounter; o[8] = u->msg2Flag;
            o[9] = u->connectionRequest; o[10] = u->msg4Flag; o[11] = u->raFailed; o[12] = u->nowBackoff;
            o[13] = 0; o[14] = 0; o[15] = 0;
        }
    }
    res->simTimeMs = time; res->nSuccess = nSuccess; res->preambleTxSum = txSum; res->delaySum = delaySum;
    res->collisionPreambles = collisionPreambles; res->totalPreambleTxop = totalPreambleTxop;
    res->captured = 1;
    clock_gettime(CLOCK_MONOTONIC, &t1);
    res->seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    free(c.ue); free(cnt); free(coll); free(phantom);
    return 0;
}
