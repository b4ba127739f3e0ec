/* ef_prof.c — tiny sampling profiler for the est-fact host code (developer tool, linked only into the -DEF_GPROF build).
 * SIGPROF every EF_PROF_US microseconds (default 1000) of process CPU time; records the interrupted PC when it lies in the program's text, else the first
 * stack word that does (the caller in our code of the libc routine that was running).  Dumped at exit as "addr count kind". */
#define _GNU_SOURCE
#include <signal.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <time.h>
#include <ucontext.h>
extern char __executable_start, etext;
#define NS (1 << 16)
static struct { uintptr_t pc; unsigned n[2]; } g_tab[NS];
static unsigned g_lost;
static void on_prof(int sig, siginfo_t *si, void *uc_) {
  (void)sig; (void)si;
  ucontext_t *uc = uc_;
  uintptr_t pc = (uintptr_t)uc->uc_mcontext.gregs[REG_RIP], lo = (uintptr_t)&__executable_start, hi = (uintptr_t)&etext;
  int kind = 0;
  if (pc < lo || pc >= hi) {
    kind = 1;
    const uintptr_t *sp = (const uintptr_t *)uc->uc_mcontext.gregs[REG_RSP];
    pc = 0;
    for (int i = 0; i < 64; ++i) if (sp[i] >= lo && sp[i] < hi) { pc = sp[i]; break; }
    if (!pc) { __atomic_fetch_add(&g_lost, 1, __ATOMIC_RELAXED); return; }
  }
  unsigned h = (unsigned)((pc * 0x9E3779B97F4A7C15ull) >> 48);
  for (int i = 0; i < 64; ++i, h = (h + 1) & (NS - 1)) {
    uintptr_t cur = __atomic_load_n(&g_tab[h].pc, __ATOMIC_RELAXED);
    if (cur == 0) { uintptr_t z = 0; if (__atomic_compare_exchange_n(&g_tab[h].pc, &z, pc, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) cur = pc; else cur = z; }
    if (cur == pc) { __atomic_fetch_add(&g_tab[h].n[kind], 1, __ATOMIC_RELAXED); return; }
  }
  __atomic_fetch_add(&g_lost, 1, __ATOMIC_RELAXED);
}
static void dump(void) {
  struct itimerval z; memset(&z, 0, sizeof z); setitimer(ITIMER_PROF, &z, NULL);
  FILE *f = fopen("ef_prof.out", "w");
  if (!f) return;
  for (int i = 0; i < NS; ++i) if (g_tab[i].pc) fprintf(f, "%lx %u %u\n", (unsigned long)(g_tab[i].pc - (uintptr_t)&__executable_start), g_tab[i].n[0], g_tab[i].n[1]);
  fprintf(f, "lost %u\n", g_lost);
  fclose(f);
}
__attribute__((constructor)) static void start(void) {
  struct sigaction sa; memset(&sa, 0, sizeof sa);
  sa.sa_sigaction = on_prof; sa.sa_flags = SA_SIGINFO | SA_RESTART;
  sigaction(SIGPROF, &sa, NULL);
  const char *ov = getenv("EF_PROF_US");
  const long us = ov && atol(ov) > 0 ? atol(ov) : 1000;
  timer_t t;
  struct sigevent ev; memset(&ev, 0, sizeof ev);
  ev.sigev_notify = SIGEV_SIGNAL; ev.sigev_signo = SIGPROF;
  if (timer_create(CLOCK_PROCESS_CPUTIME_ID, &ev, &t) == 0) {          /* high-resolution; ITIMER_PROF is tick-bound */
    struct itimerspec its = {{us / 1000000, (us % 1000000) * 1000}, {us / 1000000, (us % 1000000) * 1000}};
    timer_settime(t, 0, &its, NULL);
  } else {
    struct itimerval it = {{0, us}, {0, us}};
    setitimer(ITIMER_PROF, &it, NULL);
  }
  atexit(dump);
}
