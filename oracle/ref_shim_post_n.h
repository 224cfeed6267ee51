/*
 * ref_shim_post_n.h -- TEST INFRASTRUCTURE (oracle).  Post-shim for NOMA.c (variant N), same
 * role as ref_shim_post.h: hooks + ref_run() around the reference's own main loop (N:637-719).
 * NOMA.c has no CLI; its parameters are file-scope globals (N:41-57) set here before ref_main.
 * Draw tape: UE draws keyed (ue, ms, k) -- activeUE(user, time) N:133,142,168,187 and
 * msg2Results(user, time+1) N:460,482 through the rand() macro; resourceRequestAllocation
 * N:503,519,520 and the base-station-side draws N:284,286 (keyed (sector, ms, k), tag BS)
 * through the sed lines in build_ref.sh.
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
int ref_p_nPreamble, ref_p_backoff, ref_p_nGrantUL, ref_p_maxRarWindow, ref_p_maxMsg2TxCount,
    ref_p_accessTime, ref_p_distribution;

static ref_config  g_cfg;
static ref_result* g_res;
static int*        g_perUE;
static double*     g_gain;
static jmp_buf     g_jmp;
static int*        g_lastMs;
static unsigned short* g_cnt;
static int         g_bsLastMs[8];
static unsigned    g_bsCnt[8];
static void*       g_ueArray;
static size_t      g_ueCount;

int ref_tape_rand(int ue, int ms) {
    if (ms > g_res->lastMs) g_res->lastMs = ms;
    g_res->draws++;
    if (!g_cfg.useTape) return rand();
    if (g_lastMs[ue] != ms) { g_lastMs[ue] = ms; g_cnt[ue] = 0; }
    unsigned k = g_cnt[ue]++;
    if ((int)k + 1 > g_res->maxDrawsPerUeMs) g_res->maxDrawsPerUeMs = (int)k + 1;
    return rach_tape_rand31(g_cfg.seed, (uint32_t)g_cfg.rep, (uint32_t)ue, (uint32_t)ms, k, RACH_TAPE_TAG_UE);
}

int ref_tape_rand_bs(int sector, int ms) {
    g_res->draws++;
    if (!g_cfg.useTape) return rand();
    if (g_bsLastMs[sector] != ms) { g_bsLastMs[sector] = ms; g_bsCnt[sector] = 0; }
    unsigned k = g_bsCnt[sector]++;
    return rach_tape_rand31(g_cfg.seed, (uint32_t)g_cfg.rep, (uint32_t)sector, (uint32_t)ms, k, RACH_TAPE_TAG_BS);
}

void ref_tape_srand(unsigned seed) { (void)seed; if (!g_cfg.useTape) srand((unsigned)g_cfg.seed); }

void* ref_calloc_hook(size_t n, size_t sz) {
    void* p = calloc(n, sz);
    if (sz == sizeof(UserInfo) && n == (size_t)ref_nue) { g_ueArray = p; g_ueCount = n; }
    return p;
}

static void ref_capture(UserInfo* UE, size_t n) {
    long long txSum = 0, delaySum = 0; int nS = 0;
    for (size_t i = 0; i < n; ++i) {
        UserInfo* u = UE + i;
        if (u->RA == 1) { nS++; txSum += u->nTxPreamble; delaySum += u->timer; }
        if (g_perUE) {
            int* o = g_perUE + i * 16;
            o[0] = u->timer; o[1] = u->active; o[2] = u->txTime; o[3] = u->firstTxTime;
            o[4] = u->secondTxTime; o[5] = u->nowBackoff; o[6] = u->preamble; o[7] = u->sector;
            o[8] = u->rarWindow; o[9] = u->msg1ReTx; o[10] = u->nTxPreamble; o[11] = u->msg2;
            o[12] = u->msg3Wait; o[13] = u->RA; o[14] = u->msg3Faile; o[15] = u->RaFailed;
        }
        if (g_gain) g_gain[i] = u->channelGain;
    }
    g_res->nSuccess = nS; g_res->preambleTxSum = txSum; g_res->delaySum = delaySum;
    g_res->captured = 1;
}

void ref_free_hook(void* p) {
    if (p && p == g_ueArray) { ref_capture((UserInfo*)p, g_ueCount); g_ueArray = NULL; }
    free(p);
}

int ref_printf_hook(const char* fmt, ...) {
    if (g_cfg.echo) { va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap); }
    return 0;
}

FILE* ref_fopen_hook(const char* name, const char* mode) {
    if (g_cfg.echo >= 2) return fopen(name, mode);
    (void)name; (void)mode;
    return fopen("/dev/null", "w");
}

void ref_exit_hook(int code) { (void)code; longjmp(g_jmp, 1); }
int ref_variant(void) { return 2; }
int ref_sizeof_ue(void) { return (int)sizeof(UserInfo); }

/* perUE: nUE*16 ints (timer active txTime firstTxTime secondTxTime nowBackoff preamble sector
 * rarWindow msg1ReTx nTxPreamble msg2 msg3Wait RA msg3Faile RaFailed); geom: nUE doubles (channelGain) */
int ref_run(const ref_config* cfg, ref_result* res, int* perUE, float* geom) {
    g_cfg = *cfg; g_res = res; g_perUE = perUE; g_gain = (double*)geom;
    memset(res, 0, sizeof(*res));
    res->lastMs = -1; res->simTimeMs = -1;
    ref_nue = cfg->nUE;
    nPreamble = cfg->nPreamble; backoffIndicator = cfg->backoffIndicator; nGrantUL = cfg->nGrantUL;
    maxRarWindow = cfg->maxRarWindow; maxMsg1ReTx = cfg->maxMsg2TxCount; accessTime = cfg->accessTime;
    cellRadius = cfg->cellRadius;
    g_lastMs = (int*)malloc(sizeof(int) * (size_t)cfg->nUE);
    g_cnt = (unsigned short*)calloc((size_t)cfg->nUE, sizeof(unsigned short));
    for (int i = 0; i < cfg->nUE; ++i) g_lastMs[i] = -1;
    for (int s = 0; s < 8; ++s) { g_bsLastMs[s] = -1; g_bsCnt[s] = 0; }
    g_ueArray = NULL;
    char* argv[2] = {(char*)"ref", NULL};
    struct timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
    int rc = setjmp(g_jmp);
    if (rc == 0) ref_main(1, argv);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    res->seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    res->aborted = rc;
    free(g_lastMs); free(g_cnt);
    return rc == 1 ? -1 : 0;
}
