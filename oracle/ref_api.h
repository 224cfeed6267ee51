/*
 * ref_api.h -- TEST INFRASTRUCTURE (oracle).  ABI shared by the tape-mode builds of the
 * reference sources (oracle/_ref/libref_*.so, made by build_ref.sh) and by the C
 * restatement (oracle/rach_oracle.c -> oracle/_build/librach_oracle.so), so that tests
 * can run the same configuration through either and compare field by field.
 */
#ifndef REF_API_H
#define REF_API_H

typedef struct ref_config {
    int nUE;
    int distribution;       /* 1 = Uniform 60 s, else Beta 10 s (W encoding, W:208,241)      */
    int nPreamble, backoffIndicator, nGrantUL;
    int maxRarWindow;       /* internal value = RAR window + 1 (W:76,128)                     */
    int maxMsg2TxCount;     /* internal value = max retx - 1   (W:77,134)                     */
    int accessTime;
    float cellRadius, hBS, hUT;
    int geometry;           /* oracle only: 1 = W (2 activation draws), 0 = B                 */
    unsigned long long seed;
    int rep;
    int useTape;            /* 1 = Philox draw tape, 0 = libc rand() seeded with (unsigned)seed */
    int stopMs;             /* >0: abandon the run when a draw is requested at ms >= stopMs    */
    int echo;               /* 1 = let the reference's printf through; 2 = also write its result files */
} ref_config;

typedef struct ref_result {
    int simTimeMs, nSuccess;
    long long preambleTxSum, delaySum, failCountSum;
    long long continueFailed;
    long long collisionPreambles, totalPreambleTxop;  /* W flavour, W:62-63,625,650-652 */
    long long collisionScans, totalScans;             /* B flavour, B:41-42,334,349-351 */
    long long draws;
    int maxDrawsPerUeMs, lastMs, aborted, captured;
    int failCountsPrinted, nAccessUE;
    double averageDelay, averagePreambleTx, ratioSuccess;  /* the floats the reference prints */
    double seconds;
    /* restatement only: coverage of the rare Msg3-restart-lands-on-this-ms case */
    long long lateRestarts, lateAbsorbed;
    /* restatement of NOMA.c only: the pairing test 10*log(high)-10*log(low) > 15. (N:276) is the one floating-point
     * comparison that decides an integer outcome; pairTests = how many were evaluated, minPairMargin = the smallest
     * |difference - 15| among them (how close any decision came to flipping under a different libm) */
    long long pairTests;
    double minPairMargin;
} ref_result;

/* perUE: nUE*16 ints in rach_gpu.h ra_sim_dump_ues order (may be NULL);
 * geom: nUE*6 floats angle,x,y,distance,channelGain,sector (may be NULL). */
int ref_run(const ref_config* cfg, ref_result* res, int* perUE, float* geom);

#endif
