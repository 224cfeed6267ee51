/*
 * ref_shim_post_u0.h -- TEST INFRASTRUCTURE (oracle).  Post-shim for RandomAccessSimulator.c
 * (variant U0, the oldest simulator).  As shipped it does not compile (line 321 redeclares the
 * parameter `averageDelay`) and its arrival loop runs one element past the array (line 84,
 * `i <= activeCheck`); build_ref.sh fixes exactly those two lines, both outside the state
 * machine.  Draw tape through the rand() macro (user, time in scope at U0:160,170,187,238,250,251).
 * The reference never frees the UE array, so it is captured after ref_main returns.
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
int ref_p_nPreamble = 64, ref_p_backoff = 20, ref_p_nGrantUL, ref_p_maxRarWindow, ref_p_maxMsg2TxCount,
    ref_p_accessTime, ref_p_distribution;

static ref_config  g_cfg;
static ref_result* g_res;
static jmp_buf     g_jmp;
static int*        g_lastMs;
static unsigned short* g_cnt;
static void*       g_ueArray;
static size_t      g_ueCount;

int ref_tape_rand(int ue, int ms) {
    if (ms > g_res->lastMs) g_res->lastMs = ms;
    if (g_cfg.stopMs > 0 && ms >= g_cfg.stopMs) longjmp(g_jmp, 2);
    g_res->draws++;
    if (!g_cfg.useTape) return rand();
    if (g_lastMs[ue] != ms) { g_lastMs[ue] = ms; g_cnt[ue] = 0; }
    unsigned k = g_cnt[ue]++;
    if ((int)k + 1 > g_res->maxDrawsPerUeMs) g_res->maxDrawsPerUeMs = (int)k + 1;
    return rach_tape_rand31(g_cfg.seed, (uint32_t)g_cfg.rep, (uint32_t)ue, (uint32_t)ms, k, RACH_TAPE_TAG_UE);
}
int ref_tape_rand_bs(int s, int ms) { (void)s; (void)ms; return 0; }
void ref_tape_srand(unsigned seed) { (void)seed; if (!g_cfg.useTape) srand((unsigned)g_cfg.seed); }
void* ref_calloc_hook(size_t n, size_t sz) {
    void* p = calloc(n, sz);
    if (sz == sizeof(struct UEinfo)) { g_ueArray = p; g_ueCount = n; }
    return p;
}
void ref_free_hook(void* p) { free(p); }
int ref_printf_hook(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    if (strncmp(fmt, "Total simulation time:", 22) == 0) g_res->simTimeMs = va_arg(ap, int);
    if (g_cfg.echo) { va_list ap2; va_start(ap2, fmt); vprintf(fmt, ap2); va_end(ap2); }
    va_end(ap);
    return 0;
}
FILE* ref_fopen_hook(const char* name, const char* mode) {
    if (g_cfg.echo >= 2) return fopen(name, mode);
    (void)name; (void)mode;
    return fopen("/dev/null", "w");
}
void ref_exit_hook(int code) { (void)code; longjmp(g_jmp, 1); }
int ref_variant(void) { return 3; }
int ref_sizeof_ue(void) { return (int)sizeof(struct UEinfo); }

/* perUE: nUE*16 ints: timer active txTime preamble preambleChange rarWindow maxRarCounter
 * preambleTxCounter msg2Flag connectionRequest msg4Flag raFailed nowBackoff 0 0 0 */
int ref_run(const ref_config* cfg, ref_result* res, int* perUE, float* geom) {
    (void)geom;
    g_cfg = *cfg; g_res = res;
    memset(res, 0, sizeof(*res));
    res->lastMs = -1; res->simTimeMs = -1;
    ref_nue = cfg->nUE; ref_p_nPreamble = cfg->nPreamble; ref_p_backoff = cfg->backoffIndicator;
    g_lastMs = (int*)malloc(sizeof(int) * (size_t)cfg->nUE);
    g_cnt = (unsigned short*)calloc((size_t)cfg->nUE, sizeof(unsigned short));
    for (int i = 0; i < cfg->nUE; ++i) g_lastMs[i] = -1;
    g_ueArray = NULL;
    collisionPreambles = 0; totalPreambleTxop = 0;
    char* argv[2] = {(char*)"ref", NULL};
    struct timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
    int rc = setjmp(g_jmp);
    if (rc == 0) ref_main(1, argv);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    res->seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    res->aborted = rc;
    if (g_ueArray) {
        struct UEinfo* UE = (struct UEinfo*)g_ueArray;
        long long txSum = 0, delaySum = 0; int nS = 0;
        for (size_t i = 0; i < g_ueCount; ++i) {
            struct UEinfo* u = UE + i;
            if (u->msg4Flag == 1) { nS++; txSum += u->preambleTxCounter; delaySum += u->timer; }
            if (perUE) {
                int* o = perUE + i * 16;
                o[0] = u->timer; o[1] = u->active; o[2] = u->txTime; o[3] = u->preamble; o[4] = u->preambleChange;
                o[5] = u->rarWindow; o[6] = u->maxRarCounter; o[7] = u->preambleTxCounter; o[8] = u->msg2Flag;
                o[9] = u->connectionRequest; o[10] = u->msg4Flag; o[11] = u->raFailed; o[12] = u->nowBackoff;
                o[13] = 0; o[14] = 0; o[15] = 0;
            }
        }
        res->nSuccess = nS; res->preambleTxSum = txSum; res->delaySum = delaySum;
        res->collisionPreambles = collisionPreambles; res->totalPreambleTxop = totalPreambleTxop;
        res->captured = 1;
        free(g_ueArray); g_ueArray = NULL;
    }
    free(g_lastMs); free(g_cnt);
    return rc == 1 ? -1 : 0;
}
