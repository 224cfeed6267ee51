/*
 * rach_oracle_n.c -- TEST INFRASTRUCTURE (oracle).  CPU restatement of the sector / gain-pairing
 * simulator NOMA.c (variant N): main loop N:665-711, activeUE N:131-192,
 * preambleSectorCollisionDetection N:194-324 (+ sortUE N:90-103), msg2Results N:449-498,
 * resourceRequestAllocation N:499-546, timerIncrease N:547-553, successUEs N:554-562,
 * betaDist N:563-566 -- driven by the Philox draw tape (include/rach_tape.h): UE draws keyed
 * (ue, ms, k), the base-station draws of N:284/286 keyed (sector, ms, k) with RACH_TAPE_TAG_BS.
 *
 * cfg->geometry == 0 selects NOMA.c's alternative, non-sector collision function (N:325-447, its call is
 * commented out at N:688): one grant counter per occasion, no sector buckets, one base-station draw per pair.
 *
 * PARITY PIN: validated field by field (16 ints + channelGain of every UE) against NOMA.c itself
 * compiled in tape mode (oracle/_ref/libref_n.so, build_ref.sh) by tests/test_oracle_vs_reference.py.
 * Only tests/, smoke() and bench.py's CPU legs may load this.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "rach_tape.h"
#include "ref_api.h"

#define betaF 0.0165

typedef struct {
    int timer, active, preamble, nTxPreamble, rarWindow, msg1ReTx, msg2, msg3Wait, msg3Faile;
    int txTime, nowBackoff, firstTxTime, secondTxTime, RaFailed, RA, sector;
    int drawMs, drawK;
    double channelGain;
} nue;

typedef struct { const ref_config* cfg; ref_result* res; nue* ue; int bsMs[6]; unsigned bsK[6]; } nctx;

static int n_rand(nctx* c, int i, int ms) {
    nue* u = c->ue + i;
    if (u->drawMs != ms) { u->drawMs = ms; u->drawK = 0; }
    int k = u->drawK++;
    c->res->draws++;
    if (k + 1 > c->res->maxDrawsPerUeMs) c->res->maxDrawsPerUeMs = k + 1;
    return rach_tape_rand31(c->cfg->seed, (uint32_t)c->cfg->rep, (uint32_t)i, (uint32_t)ms, (uint32_t)k, RACH_TAPE_TAG_UE);
}

static int n_rand_bs(nctx* c, int s, int ms) {
    if (c->bsMs[s] != ms) { c->bsMs[s] = ms; c->bsK[s] = 0; }
    unsigned k = c->bsK[s]++;
    c->res->draws++;
    return rach_tape_rand31(c->cfg->seed, (uint32_t)c->cfg->rep, (uint32_t)s, (uint32_t)ms, k, RACH_TAPE_TAG_BS);
}

static float n_beta(float a, float b, float x) {            /* N:563-566 */
    float betaValue = (1 / betaF) * (pow(x, (a - 1))) * (pow((1 - x), (b - 1)));
    return betaValue;
}

static int n_align(int subTime, int accessTime) {           /* N:464-475 */
    if (subTime % accessTime == 0) return subTime + 1;
    if (subTime % accessTime == 1) return subTime;
    return subTime + (accessTime - (subTime % accessTime) + 1);
}

/* N:131-192 */
static void n_activate(nctx* c, int i, int time) {
    nue* u = c->ue + i;
    const float pi = 3.14, cellRadius = c->cfg->cellRadius;
    u->active = 1;
    u->preamble = n_rand(c, i, time) % c->cfg->nPreamble;
    u->nTxPreamble++;
    u->txTime = time + 1; u->timer = 0; u->rarWindow = 0; u->msg1ReTx = 0; u->nowBackoff = 0;
    u->firstTxTime = time + 1;
    float angle = (float)n_rand(c, i, time) / (float)(2147483647) * 2 * pi;
    if (angle >= 0 && angle < ((1. / 3.) * pi)) u->sector = 0;
    else if (angle >= ((1. / 3.) * pi) && angle < ((2. / 3.) * pi)) u->sector = 1;
    else if (angle >= ((2. / 3.) * pi) && angle < 3.14) u->sector = 2;
    else if (angle >= pi && angle < ((4. / 3.) * pi)) u->sector = 3;
    else if (angle >= ((4. / 3.) * pi) && angle < ((5. / 3.) * pi)) u->sector = 4;
    else u->sector = 5;
    float r;
    while (1) {
        r = cellRadius * sqrt((float)n_rand(c, i, time) / (float)2147483647);
        if (r > 35.0) break;
    }
    float x = r * cos(angle), y = r * sin(angle);
    double env = sqrt(x * x + y * y);                       /* float products and sum, N:183 */
    double ch_g = 0, rayleigh;
    float pathloss;
    while (ch_g < 1e-7) {                                   /* N:185-189 */
        pathloss = sqrt(1 + pow(env, 2));
        rayleigh = sqrt(-2 * log((double)n_rand(c, i, time) / (double)2147483647));
        ch_g = pow(rayleigh / pathloss, 2);
    }
    u->channelGain = ch_g;
}

typedef struct { int idx; double gain; } nsingle;

int oracle_run_n(const ref_config* cfg, ref_result* res, int* perUE, float* geom) {
    const int nUE = cfg->nUE, P = cfg->nPreamble, BI = cfg->backoffIndicator, G = cfg->nGrantUL;
    const int maxRarWindow = cfg->maxRarWindow, maxMsg1ReTx = cfg->maxMsg2TxCount, A = cfg->accessTime;
    if (nUE < 1 || P < 1 || BI < 1 || A < 1) return -1;
    memset(res, 0, sizeof(*res));
    struct timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
    nctx c; c.cfg = cfg; c.res = res;
    c.ue = (nue*)calloc((size_t)nUE, sizeof(nue));
    for (int s = 0; s < 6; ++s) { c.bsMs[s] = -1; c.bsK[s] = 0; }
    for (int i = 0; i < nUE; ++i) { c.ue[i].drawMs = -1; c.ue[i].sector = -1; }   /* initUserInfo N:104-130 */
    int* cnt = (int*)malloc(sizeof(int) * 6 * (size_t)P);
    int* who = (int*)malloc(sizeof(int) * 6 * (size_t)P);
    nsingle* tx = (nsingle*)malloc(sizeof(nsingle) * (size_t)P);
    const int maxTime = 10000;
    int activeCheck = 0, arrived = 0, nSuccess = 0, time;
    int sectorGrants[6];
    long long pairTests = 0; double minPairMargin = 1e300;
    const int nonSector = (cfg->geometry == 0);
    const int nSect = nonSector ? 1 : 6;

    for (time = 0; time < maxTime; time++) {
        res->lastMs = time;
        if (time % A == 0) {                                /* N:668-697 */
            for (int s = 0; s < 6; ++s) sectorGrants[s] = 0;
            float numBetaDist = n_beta(3, 4, (float)time / (float)maxTime);
            int accessUEs = (int)ceil((float)nUE * numBetaDist / ((float)maxTime / (float)A));
            activeCheck += accessUEs;
            if (activeCheck >= nUE) activeCheck = nUE;
            for (; arrived < activeCheck; ++arrived) n_activate(&c, arrived, time);   /* N:682-686 */

            /* ---- preambleSectorCollisionDetection, N:194-324 ---- */
            memset(cnt, 0, sizeof(int) * 6 * (size_t)P);
            for (int i = 0; i < activeCheck; ++i) {
                nue* u = c.ue + i;
                if (u->RA == 0 && u->txTime == time + 1 && u->msg2 == 0 && u->nowBackoff <= 0 && u->RaFailed == 0) {
                    int k = (nonSector ? 0 : u->sector) * P + u->preamble;
                    if (cnt[k]++ == 0) who[k] = i;
                }
            }
            for (int s = 0; s < nSect; ++s) {
                int count = 0;
                for (int p = 0; p < P; ++p)                 /* singles in preamble order, N:243-249 */
                    if (cnt[s * P + p] == 1) { tx[count].idx = who[s * P + p]; tx[count].gain = c.ue[who[s * P + p]].channelGain; count++; }
                if (count == 0) continue;
                if (count <= G) {                           /* N:252-260; N:377-384 grants msg2 unconditionally */
                    for (int i = 0; i < count; ++i) {
                        if (sectorGrants[s] < G) { sectorGrants[s]++; c.ue[tx[i].idx].msg2 = 1; }
                        else if (nonSector) c.ue[tx[i].idx].msg2 = 1;
                    }
                } else {
                    /* sortUE N:90-103: bubble sort ascending by gain == a stable sort */
                    for (int i = 1; i < count; ++i) {
                        nsingle t = tx[i]; int j = i - 1;
                        while (j >= 0 && t.gain < tx[j].gain) { tx[j + 1] = tx[j]; --j; }
                        tx[j + 1] = t;
                    }
                    int pair = 0;
                    for (int i = 0; i < count - 1; ++i) {   /* N:268-298 */
                        for (int j = 1; j < count; ++j) {
                            int rx0 = tx[i].idx, rx1 = tx[j].idx;
                            double low = tx[i].gain, high = tx[j].gain;
                            if (rx0 != -1 && rx1 != -1) {
                                double margin = fabs(10 * log(high) - 10 * log(low) - 15.);
                                pairTests++;
                                if (margin < minPairMargin) minPairMargin = margin;
                            }
                            if (rx0 != -1 && rx1 != -1 && 10 * log(high) - 10 * log(low) > 15.) {
                                pair += 2;
                                tx[i].idx = -1; tx[j].idx = -1;
                                if (sectorGrants[s] < G) {
                                    sectorGrants[s]++;
                                    double p = (double)n_rand_bs(&c, s, time) / (double)2147483647;
                                    if (p < 0.3) {
                                        if (nonSector) c.ue[rx0].msg2 = 1;                 /* N:413-415 */
                                        else {
                                            int randomUE = n_rand_bs(&c, s, time) % 2;      /* N:286-287 */
                                            c.ue[randomUE ? rx1 : rx0].msg2 = 1;
                                        }
                                    } else { c.ue[rx0].msg2 = 1; c.ue[rx1].msg2 = 1; }
                                }
                                break;
                            }
                        }
                    }
                    if (count - pair > 0)                   /* N:299-307 */
                        for (int i = 0; i < count; ++i)
                            if (tx[i].idx != -1 && sectorGrants[s] < G) { sectorGrants[s]++; c.ue[tx[i].idx].msg2 = 1; }
                }
            }
            /* ---- msg2Results(UE, time+1), N:691-696, 449-498 ---- */
            for (int i = 0; i < activeCheck; ++i) {
                nue* u = c.ue + i;
                if (!(u->nowBackoff <= 0 && u->txTime == time + 1 && u->active == 1 && u->RA == 0 && u->RaFailed == 0)) continue;
                if (u->msg2 == 0 && u->active == 1) {
                    u->rarWindow = 5;
                    u->txTime += 3;
                    if (u->rarWindow >= maxRarWindow) {
                        u->nTxPreamble++; u->rarWindow = 0; u->msg1ReTx++;
                        int tmp = n_rand(&c, i, time + 1) % BI;
                        u->txTime = n_align(u->txTime + tmp, A);
                        u->nowBackoff = u->txTime - (time + 1) - 1;
                        u->secondTxTime = u->txTime;
                        if (u->msg1ReTx >= maxMsg1ReTx) {
                            u->preamble = n_rand(&c, i, time + 1) % P;
                            u->RaFailed++; u->nTxPreamble = 0; u->rarWindow = 0; u->msg1ReTx = 0; u->timer = 0;
                        }
                    }
                } else if (u->msg2 == 1) {
                    u->active = 2; u->txTime += 10; u->secondTxTime = u->txTime; u->msg3Wait = 0;
                }
            }
        }
        /* ---- resourceRequestAllocation, N:699, 499-546 ---- */
        for (int i = 0; i < activeCheck; ++i) {
            nue* u = c.ue + i;
            if (!(u->txTime == time && u->msg2 == 1 && u->active == 2 && u->RaFailed == 0)) continue;
            if (u->msg3Wait <= 48) {
                int r = n_rand(&c, i, time);
                if (rach_msg3_success(r)) { u->active = 0; u->RA = 1; u->timer = u->timer + 6; nSuccess++; }
                else { u->txTime += 49; u->msg3Wait = 49; }
            } else {
                u->RA = 0; u->msg3Faile++; u->active = 1; u->msg2 = 0;
                u->preamble = n_rand(&c, i, time) % P;
                int tmp = n_rand(&c, i, time) % BI;
                u->txTime = n_align(u->txTime + tmp, A);
                u->secondTxTime = u->txTime;
                u->nowBackoff = u->txTime - time - 1;
                u->rarWindow = 0; u->nTxPreamble = 0; u->msg1ReTx = 0; u->timer = 0;
            }
        }
        for (int i = 0; i < activeCheck; ++i) {             /* N:702-706 */
            nue* u = c.ue + i;
            if (u->active > 0 && u->RA == 0 && u->RaFailed == 0) { u->timer++; if (u->nowBackoff > 0) u->nowBackoff--; }
        }
        if (nSuccess == nUE) break;                         /* N:707-710 */
    }

    long long txSum = 0, delaySum = 0;
    for (int i = 0; i < nUE; ++i) {
        nue* u = c.ue + i;
        if (u->RA == 1) { txSum += u->nTxPreamble; delaySum += u->timer; }
        if (perUE) {
            int* o = perUE + (size_t)i * 16;
            o[0] = u->timer; o[1] = u->active; o[2] = u->txTime; o[3] = u->firstTxTime; o[4] = u->secondTxTime;
            o[5] = u->nowBackoff; o[6] = u->preamble; o[7] = u->sector; o[8] = u->rarWindow; o[9] = u->msg1ReTx;
            o[10] = u->nTxPreamble; o[11] = u->msg2; o[12] = u->msg3Wait; o[13] = u->RA; o[14] = u->msg3Faile;
            o[15] = u->RaFailed;
        }
        if (geom) ((double*)geom)[i] = u->channelGain;
    }
    res->simTimeMs = time; res->nSuccess = nSuccess; res->preambleTxSum = txSum; res->delaySum = delaySum;
    res->captured = 1; res->pairTests = pairTests; res->minPairMargin = minPairMargin;
    clock_gettime(CLOCK_MONOTONIC, &t1);
    res->seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    free(c.ue); free(cnt); free(who); free(tx);
    return 0;
}
