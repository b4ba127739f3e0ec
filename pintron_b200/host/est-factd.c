/* est-factd — the resident GPU server behind est-fact.
 *
 * pintron.py starts one est-fact process per gene under `ulimit -t .. && ulimit -v ..` (dist-scripts/pintron.py:878-884,
 * 207-213).  A CUDA context costs 0.5-4 s to create and reserves far more address space than that ulimit allows, so the
 * GPU side lives here instead: est-factd holds the contexts of the GPUs of the box, the loaded kernels and the pinned
 * lane segments; every est-fact process is a CUDA-free client that sends its genome, receives lanes (memfd segments
 * passed over the socket) and has its DP batches merged with everyone else's by the engine (csrc/pc_engine.cu).
 *
 *   est-factd [--socket PATH] [--devices 0,1,..] [--idle-timeout SECONDS] [--segment-mb N] [--foreground]
 *
 * One thread per connection; a connection = one session (one genome on one GPU).  When the client goes away — BYE, exit
 * or crash — the session's lanes and device memory are released.  With --idle-timeout N the server exits after N seconds
 * without a session (est-fact starts a new one on demand); 0 = stay.
 */
#define _GNU_SOURCE
#include "engine_client.h"
#include <errno.h>
#include <fcntl.h>
#include <pthread.h>
#include <signal.h>
#include <stdatomic.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/resource.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <time.h>
#include <unistd.h>

#define MIN2(a, b) ((a) < (b) ? (a) : (b))
static pc_engine *g_eng;
static _Atomic int g_sessions;
static _Atomic long g_last_activity;
static _Atomic int g_stop;
static char g_path[256];

static void say(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
#include <stdarg.h>
static void say(const char *fmt, ...) {
  char t[32];
  time_t now = time(NULL);
  struct tm tm;
  localtime_r(&now, &tm);
  strftime(t, sizeof t, "%H:%M:%S", &tm);
  fprintf(stderr, "[est-factd %s] ", t);
  va_list ap;
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  va_end(ap);
  fputc('\n', stderr);
  fflush(stderr);
}

static void send_error(int s, const char *msg) { efd_send(s, EFD_ERROR, msg, strlen(msg), NULL, 0); }

static void *serve(void *arg) {
  const int s = (int)(intptr_t)arg;
  uint32_t type = 0; void *pl = NULL; size_t plen = 0;
  uint32_t session = 0;
  int gpu = -1;
  if (efd_recv(s, &type, &pl, &plen, NULL, 0, NULL)) { close(s); return NULL; }
  if (type == EFD_SHUTDOWN) { free(pl); close(s); atomic_store(&g_stop, 1); return NULL; }
  if (type != EFD_HELLO || plen < sizeof(efd_hello)) { send_error(s, "expected HELLO"); free(pl); close(s); return NULL; }
  efd_hello h;
  memcpy(&h, pl, sizeof h);
  if (plen != sizeof h + h.genome_len || h.nlanes < 1 || h.nlanes > PCE_MAX_SESSION_LANES) { send_error(s, "malformed HELLO"); free(pl); close(s); return NULL; }
  pc_session_req r = {h.gpu, (const char *)pl + sizeof h, (size_t)h.genome_len, h.word_len, h.depth_rate, h.nlanes, h.arena_cap, h.var_cap, h.jobs_cap};
  pc_session_info info;
  atomic_fetch_add(&g_sessions, 1);
  struct timespec o0, o1;
  clock_gettime(CLOCK_MONOTONIC, &o0);
  const int rc = pc_engine_open(g_eng, &r, &info);
  clock_gettime(CLOCK_MONOTONIC, &o1);
  free(pl); pl = NULL;
  if (rc) {
    char msg[600];
    snprintf(msg, sizeof msg, "cannot open a session: %s", pc_last_error());
    send_error(s, msg);
    goto out;
  }
  session = info.session; gpu = info.gpu;
  if (h.timers) pc_engine_enable_timers(g_eng, 1);
  {
    efd_hello_ok ok;
    memset(&ok, 0, sizeof ok);
    ok.session = info.session; ok.gpu = info.gpu; ok.nlanes = info.nlanes;
    snprintf(ok.backend, sizeof ok.backend, "%s", pc_engine_backend());
    ok.nsegs = pc_engine_segment_count(g_eng, gpu);
    int fds[PCE_MAX_SEGMENTS];
    for (int i = 0; i < ok.nsegs; ++i) { size_t b = 0; fds[i] = pc_engine_segment_fd(g_eng, gpu, i, &b); ok.seg_bytes[i] = b; }
    memcpy(ok.lane, info.lane, sizeof(uint32_t) * (size_t)info.nlanes);
    if (efd_send(s, EFD_HELLO_OK, &ok, sizeof ok, fds, ok.nsegs)) goto out;
  }
  say("session %u opened on gpu %d: %d lanes, genome %llu bp (genome index + lanes in %.1f ms)", session, gpu, info.nlanes, (unsigned long long)h.genome_len,
      1e3 * (double)(o1.tv_sec - o0.tv_sec) + 1e-6 * (double)(o1.tv_nsec - o0.tv_nsec));
  for (;;) {
    if (efd_recv(s, &type, &pl, &plen, NULL, 0, NULL)) break;            /* EOF: the client is gone */
    if (type == EFD_BYE) {
      pc_session_stats st;
      clock_gettime(CLOCK_MONOTONIC, &o0);
      pc_engine_close(g_eng, session, &st);
      clock_gettime(CLOCK_MONOTONIC, &o1);
      say("session %u closed in %.1f ms: %llu batches (%llu lanes merged), %llu jobs, %llu launches, engine busy %.3f s", session,
          1e3 * (double)(o1.tv_sec - o0.tv_sec) + 1e-6 * (double)(o1.tv_nsec - o0.tv_nsec),
          (unsigned long long)st.batches, (unsigned long long)st.lanes_merged, (unsigned long long)st.jobs, (unsigned long long)st.launches, st.busy_s);
      session = 0;
      efd_send(s, EFD_STATS, &st, sizeof st, NULL, 0);
      free(pl); pl = NULL;
      break;
    }
    if (type == EFD_RESIZE && plen == sizeof(efd_resize)) {
      efd_resize q;
      memcpy(&q, pl, sizeof q);
      free(pl); pl = NULL;
      if (pc_engine_resize_lane(g_eng, session, q.lane, q.arena_cap, q.jobs_cap, q.var_cap, q.keep_arena, q.keep_jobs)) {
        char msg[600];
        snprintf(msg, sizeof msg, "cannot resize lane %u: %s", q.lane, pc_last_error());
        send_error(s, msg);
        continue;
      }
      efd_resize_ok ok;
      memset(&ok, 0, sizeof ok);
      ok.nsegs = pc_engine_segment_count(g_eng, gpu);
      int fds[PCE_MAX_SEGMENTS], nf = 0;
      for (int i = 0; i < ok.nsegs; ++i) {
        size_t b = 0;
        const int fd = pc_engine_segment_fd(g_eng, gpu, i, &b);
        ok.seg_bytes[i] = b;
        if ((uint32_t)i >= q.have_segs) fds[nf++] = fd;
      }
      ok.new_fds = nf;
      if (efd_send(s, EFD_RESIZE_OK, &ok, sizeof ok, fds, nf)) break;
      continue;
    }
    free(pl); pl = NULL;
    send_error(s, "unknown request");
  }
out:
  if (session) { pc_engine_close(g_eng, session, NULL); say("session %u: client went away, released", session); }
  close(s);
  atomic_store(&g_last_activity, (long)time(NULL));
  atomic_fetch_sub(&g_sessions, 1);
  return NULL;
}

static void on_term(int sig) { (void)sig; atomic_store(&g_stop, 1); }

int main(int argc, char **argv) {
  int devices[16], ndev = 0, idle = 0;
  size_t seg_mb = 0;
  bool stop = false;
  efd_default_socket(g_path, sizeof g_path);
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--socket") && i + 1 < argc) snprintf(g_path, sizeof g_path, "%s", argv[++i]);
    else if (!strcmp(argv[i], "--idle-timeout") && i + 1 < argc) idle = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--segment-mb") && i + 1 < argc) seg_mb = (size_t)atol(argv[++i]);
    else if (!strcmp(argv[i], "--devices") && i + 1 < argc) {
      char *sv = NULL, *str = strdup(argv[++i]);
      for (char *t = strtok_r(str, ",", &sv); t && ndev < 16; t = strtok_r(NULL, ",", &sv)) devices[ndev++] = atoi(t);
      free(str);
    } else if (!strcmp(argv[i], "--foreground")) {
    } else if (!strcmp(argv[i], "--stop")) stop = true;
    else {
      fprintf(stderr, "usage: est-factd [--socket PATH] [--devices 0,1,..] [--idle-timeout S] [--segment-mb N] [--foreground] [--stop]\n");
      return strcmp(argv[i], "--help") ? 1 : 0;
    }
  }
  if (stop) {                                                  /* ask a running server to exit */
    struct sockaddr_un a; memset(&a, 0, sizeof a); a.sun_family = AF_UNIX; memcpy(a.sun_path, g_path, MIN2(sizeof a.sun_path - 1, strlen(g_path)));
    int s = socket(AF_UNIX, SOCK_STREAM, 0);
    if (s < 0 || connect(s, (struct sockaddr *)&a, sizeof a)) { fprintf(stderr, "est-factd: nothing listening on %s\n", g_path); return 1; }
    efd_send(s, EFD_SHUTDOWN, NULL, 0, NULL, 0);
    close(s);
    return 0;
  }
  /* started from inside a job's ulimit: lift what we are allowed to (the limits are the job's, not the server's) */
  {
    struct rlimit rl;
    if (getrlimit(RLIMIT_AS, &rl) == 0 && rl.rlim_cur != RLIM_INFINITY) {
      struct rlimit want = {RLIM_INFINITY, RLIM_INFINITY};
      if (setrlimit(RLIMIT_AS, &want) != 0) { want.rlim_cur = want.rlim_max = rl.rlim_max; setrlimit(RLIMIT_AS, &want); }
    }
    if (getrlimit(RLIMIT_CPU, &rl) == 0 && rl.rlim_cur != RLIM_INFINITY) {
      struct rlimit want = {RLIM_INFINITY, RLIM_INFINITY};
      if (setrlimit(RLIMIT_CPU, &want) != 0) { want.rlim_cur = want.rlim_max = rl.rlim_max; setrlimit(RLIMIT_CPU, &want); }
    }
  }
  signal(SIGPIPE, SIG_IGN);
  signal(SIGTERM, on_term);
  signal(SIGINT, on_term);
  /* load every kernel with the context (one thread), not lazily under the first clients' batches */
  setenv("CUDA_MODULE_LOADING", "EAGER", 0);

  /* claim the socket first: of several servers started at the same moment only one goes on to create contexts */
  int ls = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
  struct sockaddr_un a;
  memset(&a, 0, sizeof a);
  a.sun_family = AF_UNIX;
  memcpy(a.sun_path, g_path, MIN2(sizeof a.sun_path - 1, strlen(g_path)));
  {
    int probe = socket(AF_UNIX, SOCK_STREAM, 0);
    if (probe >= 0 && connect(probe, (struct sockaddr *)&a, sizeof a) == 0) { close(probe); say("another est-factd already serves %s", g_path); return 0; }
    if (probe >= 0) close(probe);
    unlink(g_path);                                             /* stale socket file of a dead server */
  }
  char lock[300];
  snprintf(lock, sizeof lock, "%s.lock", g_path);
  int lfd = open(lock, O_CREAT | O_RDWR | O_CLOEXEC, 0600);
  if (lfd < 0 || lockf(lfd, F_TLOCK, 0) != 0) { say("another est-factd is starting on %s", g_path); return 0; }

  if (ndev == 0) {
    int n = pc_device_count();
    if (n <= 0) { say("no CUDA device: %s", pc_last_error()); return 1; }
    for (int i = 0; i < n && i < 16; ++i) devices[ndev++] = i;
  }
  const double t0 = (double)clock() / CLOCKS_PER_SEC;
  struct timespec w0, w1;
  clock_gettime(CLOCK_MONOTONIC, &w0);
  g_eng = pc_engine_create(devices, ndev, seg_mb << 20);
  if (!g_eng) { say("cannot start the engine: %s", pc_last_error()); return 1; }
  clock_gettime(CLOCK_MONOTONIC, &w1);
  (void)t0;
  mode_t old = umask(0077);
  if (ls < 0 || bind(ls, (struct sockaddr *)&a, sizeof a) != 0 || listen(ls, 128) != 0) { say("cannot listen on %s: %s", g_path, strerror(errno)); return 1; }
  umask(old);
  say("serving %d GPU(s) on %s (engine up in %.3f s, idle timeout %d s)", ndev, g_path,
      (double)(w1.tv_sec - w0.tv_sec) + 1e-9 * (double)(w1.tv_nsec - w0.tv_nsec), idle);
  atomic_store(&g_last_activity, (long)time(NULL));
  while (!atomic_load(&g_stop)) {
    fd_set rf;
    FD_ZERO(&rf);
    FD_SET(ls, &rf);
    struct timeval tv = {0, 200000};
    const int k = select(ls + 1, &rf, NULL, NULL, &tv);
    if (k > 0) {
      const int s = accept4(ls, NULL, NULL, SOCK_CLOEXEC);
      if (s < 0) continue;
      atomic_store(&g_last_activity, (long)time(NULL));
      pthread_t th;
      pthread_attr_t at;
      pthread_attr_init(&at);
      pthread_attr_setdetachstate(&at, PTHREAD_CREATE_DETACHED);
      if (pthread_create(&th, &at, serve, (void *)(intptr_t)s)) close(s);
      pthread_attr_destroy(&at);
    } else if (idle > 0 && atomic_load(&g_sessions) == 0 && (long)time(NULL) - atomic_load(&g_last_activity) > idle) {
      say("idle for %d s: leaving", idle);
      break;
    }
  }
  unlink(g_path);
  unlink(lock);
  say("stopped");
  _exit(0);            /* skip the CUDA runtime's exit-time tear-down: nothing of ours is left to flush */
}
