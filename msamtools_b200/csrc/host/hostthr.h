/*
 * hostthr.h -- the host side's worker threads step back for its I/O threads.
 *
 * `filter | profile` moves gigabytes through a pipe whose two ends (one writing thread, one reading thread) are woken up once per
 * megabyte; with 2 x 16 inflate / pack workers runnable on the same cores every wake-up queues behind a time slice and the pipe
 * runs at half its rate (1.84 -> 0.94 GB/s measured with 16 busy threads on 8 cores).  Spawned workers therefore lower their own
 * priority (nice +10, which needs no privilege and changes nothing on an idle machine); the threads that feed pipes and the GPU
 * keep the process's.
 */
#ifndef MSG_HOSTTHR_H
#define MSG_HOSTTHR_H
#ifdef __linux__
#include <sys/resource.h>
#include <sys/syscall.h>
#include <unistd.h>
static inline void worker_step_back(void) { (void)setpriority(PRIO_PROCESS, (id_t)syscall(SYS_gettid), 10); }
#else
static inline void worker_step_back(void) { }
#endif
#endif
