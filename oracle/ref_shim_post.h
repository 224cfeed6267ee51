/*
 * ref_shim_post.h -- TEST INFRASTRUCTURE (oracle).  Not part of the product path.
 *
 * Appended (same translation unit) after the reference text, so `struct UEinfo`, the
 * reference's file-scope counters (RandomAccessWithNOMA.c:62-63) and `ref_main` are
 * visible.  Implements the hooks declared in ref_shim_pre.h and exports
 *
 *     int ref_run(const ref_config*, ref_result*, int* perUE, float* geom)
 *
 * which runs ONE (seed, nUE) point of the reference's own main loop.
 *
 * REF_VARIANT: 0 = RandomAccessWithNOMA.c (W), 1 = RandomAccessSimulatorBeta.c (B).
 */
#undef rand
#undef srand
#undef calloc
#undef free
#undef printf
#undef fopen
#undef exit
#undef mkdir
#undef main

#include "rach_tape.h"
#include "ref_api.h"

int ref_nue = 10000;
int ref_p_nPreamble = 54, ref_p_backoff = 20, ref_p_nGrantUL = 54, ref_p_maxRarWindow = 6,
    ref_p_maxMsg2TxCount = 9, ref_p_accessTime = 5, ref_p_distribution = 1;

static ref_config  g_cfg;
static ref_result* g_res;
static int*        g_perUE;
static float*      g_geom;
static jmp_buf     g_jmp;
static int*        g_lastMs;
static unsigned short* g_cnt;
static void*       g_ueArray;
static size_t      g_ueCount;
static FILE*       g_devnull;

int ref_tape_rand(int ue, int ms) {
    if (ms > g_res->lastMs) g_res->lastMs = ms;
    if (g_cfg.stopMs > 0 && ms >= g_cfg.stopMs) longjmp(g_jmp, 2);
    g_res->draws++;
    if (!g_cfg.useTape) return rand();
    if (g_lastMs[ue] != ms) { g_lastMs[ue] = ms; g_cnt[ue] = 0; }
    unsigned k = g_cnt[ue]++;
    if ((int)k > g_res->maxDrawsPerUeMs - 1) g_res->maxDrawsPerUeMs = (int)k + 1;
    return rach_tape_rand31(g_cfg.seed, (uint32_t)g_cfg.rep, (uint32_t)ue, (uint32_t)ms, k,
                            RACH_TAPE_TAG_UE);
}

void ref_tape_srand(unsigned seed) {
    (void)seed;
    if (!g_cfg.useTape) srand((unsigned)g_cfg.seed);
}

void* ref_calloc_hook(size_t n, size_t sz) {
    void* p = calloc(n, sz);
    if (sz == sizeof(struct UEinfo)) { g_ueArray = p; g_ueCount = n; }
    return p;
}

static void ref_capture(struct UEinfo* UE, size_t n) {
    long long txSum = 0, delaySum = 0, failSum = 0; int nS = 0;
    for (size_t i = 0; i < n; ++i) {
        struct UEinfo* u = UE + i;
        if (u->msg4Flag == 1) {
            nS++; txSum += u->preambleTxCounter; delaySum += u->timer;
#if REF_VARIANT == 0
            failSum += u->failCount;
#endif
        }
        if (g_perUE) {
            int* o = g_perUE + i * 16;
            o[0] = u->timer; o[1] = u->active; o[2] = u->txTime; o[3] = u->firstTxTime;
            o[4] = u->secondTxTime; o[5] = u->nowBackoff; o[6] = u->preamble;
            o[7] = u->preambleChange; o[8] = u->rarWindow; o[9] = u->maxRarCounter;
            o[10] = u->preambleTxCounter; o[11] = u->msg2Flag; o[12] = u->connectionRequest;
            o[13] = u->msg4Flag;
#if REF_VARIANT == 0
            o[14] = u->failCount; o[15] = (u->active == -1) ? -1 : u->sector;
#else
            o[14] = 0; o[15] = -1;
#endif
        }
#if REF_VARIANT == 0
        if (g_geom) {
            float* o = g_geom + i * 6;
            o[0] = u->angle; o[1] = u->xCoordinate; o[2] = u->yCoordinate; o[3] = u->distance;
            o[4] = u->channelGain; o[5] = (float)u->sector;
        }
#endif
    }
    g_res->nSuccess = nS; g_res->preambleTxSum = txSum; g_res->delaySum = delaySum;
    g_res->failCountSum = failSum;
#if REF_VARIANT == 0
    g_res->collisionPreambles = collisionPreambles;
    g_res->totalPreambleTxop = totalPreambleTxop;
#else
    g_res->collisionScans = collisionPreambles;
    g_res->totalScans = totalPreambleTxop;
#endif
    g_res->captured = 1;
}

void ref_free_hook(void* p) {
    if (p && p == g_ueArray) { ref_capture((struct UEinfo*)p, g_ueCount); g_ueArray = NULL; }
    free(p);
}

int ref_printf_hook(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    if (strncmp(fmt, "Total simulation time:", 22) == 0)      g_res->simTimeMs = va_arg(ap, int);
    else if (strncmp(fmt, "Number of falied UEs:", 21) == 0)  g_res->continueFailed = va_arg(ap, int);
    else if (strncmp(fmt, "Fail Counts:", 12) == 0)           g_res->failCountsPrinted = va_arg(ap, int);
    else if (strncmp(fmt, "Number of RA try UEs per Subframe:", 34) == 0) g_res->nAccessUE = va_arg(ap, int);
    else if (strncmp(fmt, "Average delay:", 14) == 0)         g_res->averageDelay = va_arg(ap, double);
    else if (strncmp(fmt, "Average preamble tx count:", 26) == 0) g_res->averagePreambleTx = va_arg(ap, double);
    else if (strncmp(fmt, "Success ratio:", 14) == 0)         g_res->ratioSuccess = va_arg(ap, double);
    if (g_cfg.echo) { va_list ap2; va_start(ap2, fmt); vprintf(fmt, ap2); va_end(ap2); }
    va_end(ap);
    return 0;
}

FILE* ref_fopen_hook(const char* name, const char* mode) {
    if (g_cfg.echo >= 2) return fopen(name, mode);     /* format tests: real files in the cwd */
    (void)name; (void)mode;
    return fopen("/dev/null", "w");
}

void ref_exit_hook(int code) { (void)code; longjmp(g_jmp, 1); }

int ref_variant(void) { return REF_VARIANT; }
int ref_sizeof_ue(void) { return (int)sizeof(struct UEinfo); }

int ref_run(const ref_config* cfg, ref_result* res, int* perUE, float* geom) {
    g_cfg = *cfg; g_res = res; g_perUE = perUE; g_geom = geom;
    memset(res, 0, sizeof(*res));
    res->lastMs = -1; res->simTimeMs = -1;
    ref_nue = cfg->nUE;
    g_lastMs = (int*)malloc(sizeof(int) * (size_t)cfg->nUE);
    g_cnt = (unsigned short*)calloc((size_t)cfg->nUE, sizeof(unsigned short));
    for (int i = 0; i < cfg->nUE; ++i) g_lastMs[i] = -1;
    g_ueArray = NULL;
    collisionPreambles = 0; totalPreambleTxop = 0;

    char a[12][32]; char* argv[32]; int argc = 0;
    argv[argc++] = (char*)"ref";
#if REF_VARIANT == 0
    /* RandomAccessWithNOMA.c:90-158: the flags the parser really accepts */
    int k = 0;
#define REF_ARG(flag, fmtspec, val) do { argv[argc++] = (char*)(flag); \
        snprintf(a[k], sizeof a[k], fmtspec, val); argv[argc++] = a[k++]; } while (0)
    REF_ARG("-t", "%d", 1);
    REF_ARG("-d", "%d", cfg->distribution == 1 ? 1 : 0);
    REF_ARG("-p", "%d", cfg->nPreamble);
    REF_ARG("-b", "%d", cfg->backoffIndicator);
    REF_ARG("-g", "%d", cfg->nGrantUL);
    REF_ARG("-rc", "%d", cfg->maxRarWindow - 1);
    REF_ARG("-mrc", "%d", cfg->maxMsg2TxCount + 1);
    REF_ARG("-s", "%d", cfg->accessTime);
    REF_ARG("-c", "%.9g", (double)cfg->cellRadius);
    REF_ARG("-bs", "%.9g", (double)cfg->hBS);
    REF_ARG("-ut", "%.9g", (double)cfg->hUT);
#undef REF_ARG
#else
    (void)a;
    ref_p_nPreamble = cfg->nPreamble; ref_p_backoff = cfg->backoffIndicator;
    ref_p_nGrantUL = cfg->nGrantUL; ref_p_maxRarWindow = cfg->maxRarWindow;
    ref_p_maxMsg2TxCount = cfg->maxMsg2TxCount; ref_p_accessTime = cfg->accessTime;
    ref_p_distribution = cfg->distribution == 1 ? 0 : 1;   /* B:57: 0 Uniform, 1 Beta */
#endif
    argv[argc] = NULL;

    struct timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
    int rc = setjmp(g_jmp);
    if (rc == 0) { ref_main(argc, argv); }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    res->seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    res->aborted = rc;
    if (rc == 2 && g_ueArray) {          /* stopped at stopMs: capture what there is */
        ref_capture((struct UEinfo*)g_ueArray, g_ueCount);
        free(g_ueArray); g_ueArray = NULL;
    }
    free(g_lastMs); free(g_cnt);
    return rc == 1 ? -1 : 0;
}
