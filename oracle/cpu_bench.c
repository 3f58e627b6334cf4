/*
 * oracle/cpu_bench.c -- CPU baseline runner: one firpfbch2 analyser object per host
 * thread, each over its own slice of the stimulus.  Used only by bench.py's
 * cpu_baseline / --impl reference legs.  TEST INFRASTRUCTURE ONLY (see yagi_oracle.h).
 *
 * yagi itself is single-threaded (no threads anywhere); "one
 * channelizer object per core" is how BASELINE.json's north_star asks the CPU path
 * to be timed.
 */
#define _GNU_SOURCE
#include "yagi_oracle.h"

#include <pthread.h>
#include <sched.h>
#include <stdlib.h>
#include <time.h>

typedef struct {
    orc_firpfbch2* q;
    const ocf32*   x;
    ocf32*         y;
    ocf32*         y_local;
    size_t         n_frames;
    uint32_t       passes;
    uint32_t       M;
    int            cpu;
    pthread_barrier_t* start;
    double         seconds;
} worker_t;

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void* worker(void* arg)
{
    worker_t* w = (worker_t*)arg;
    if (w->cpu >= 0) {
        cpu_set_t set;
        CPU_ZERO(&set);
        CPU_SET(w->cpu, &set);
        pthread_setaffinity_np(pthread_self(), sizeof(set), &set);   /* best effort */
    }
    ocf32* y = w->y ? w->y : w->y_local;
    pthread_barrier_wait(w->start);
    const double t0 = now_s();
    for (uint32_t p = 0; p < w->passes; p++) {
        if (w->y) {
            orc_firpfbch2_crcf_execute_block(w->q, w->x, w->n_frames, y);
        } else {
            /* no output buffer supplied: reuse one frame of scratch per call */
            const size_t M2 = w->M / 2;
            for (size_t k = 0; k < w->n_frames; k++)
                orc_firpfbch2_crcf_execute(w->q, w->x + k * M2, y);
        }
    }
    w->seconds = now_s() - t0;
    return NULL;
}

double orc_bench_firpfbch2_analysis(uint32_t M, uint32_t m, float as,
                                    const ocf32* x, size_t n_per_thread,
                                    uint32_t n_threads, uint32_t passes, ocf32* y)
{
    if (n_threads == 0 || M < 2 || (M & 1)) return -1.0;
    const size_t n_frames = n_per_thread / (M / 2);
    worker_t* ws = (worker_t*)calloc(n_threads, sizeof(worker_t));
    pthread_t* th = (pthread_t*)calloc(n_threads, sizeof(pthread_t));
    pthread_barrier_t start;
    pthread_barrier_init(&start, NULL, n_threads);

    cpu_set_t avail;
    CPU_ZERO(&avail);
    int have_aff = sched_getaffinity(0, sizeof(avail), &avail) == 0;
    int next_cpu = 0;

    for (uint32_t t = 0; t < n_threads; t++) {
        if (orc_firpfbch2_crcf_create_kaiser(ORC_ANALYZER, M, m, as, &ws[t].q)) return -1.0;
        ws[t].x = x + (size_t)t * n_per_thread;
        ws[t].y = y ? y + (size_t)t * 2 * n_per_thread : NULL;
        ws[t].y_local = y ? NULL : (ocf32*)malloc(sizeof(ocf32) * M);
        ws[t].n_frames = n_frames;
        ws[t].passes = passes;
        ws[t].M = M;
        ws[t].start = &start;
        ws[t].cpu = -1;
        if (have_aff) {
            while (next_cpu < CPU_SETSIZE && !CPU_ISSET(next_cpu, &avail)) next_cpu++;
            if (next_cpu < CPU_SETSIZE) ws[t].cpu = next_cpu++;
        }
    }
    for (uint32_t t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, worker, &ws[t]);
    double worst = 0.0;
    for (uint32_t t = 0; t < n_threads; t++) {
        pthread_join(th[t], NULL);
        if (ws[t].seconds > worst) worst = ws[t].seconds;
        orc_firpfbch2_crcf_destroy(ws[t].q);
        free(ws[t].y_local);
    }
    pthread_barrier_destroy(&start);
    free(ws);
    free(th);
    return worst;
}
