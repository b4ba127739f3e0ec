/* engine_client.c — est-fact's connection to the batch engine: in-process, or the resident server est-factd.
 *
 * The reference est-fact is one self-contained process per gene (dist-scripts/pintron.py:878-884).  Ours keeps that
 * contract — same argv, same files, same exit codes — but the GPU side may live in a server that outlasts the job:
 * the CUDA context, the pinned lane segments and the loaded kernels are then paid once per box instead of once per
 * gene, and this process never initialises CUDA (so pintron.py's default `ulimit -v` holds, pintron.py:207-213). */
#define _GNU_SOURCE
#include "engine_client.h"
#include <errno.h>
#include <fcntl.h>
#include <poll.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/resource.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

/* ---- wire helpers ------------------------------------------------------------------------------------------------ */
static int write_all(int fd, const void *p, size_t n) {
  const char *c = p;
  while (n) {
    ssize_t k = send(fd, c, n, MSG_NOSIGNAL);
    if (k < 0) { if (errno == EINTR) continue; return -1; }
    c += k; n -= (size_t)k;
  }
  return 0;
}
static int read_all(int fd, void *p, size_t n) {
  char *c = p;
  while (n) {
    ssize_t k = recv(fd, c, n, 0);
    if (k < 0) { if (errno == EINTR) continue; return -1; }
    if (k == 0) return -1;
    c += k; n -= (size_t)k;
  }
  return 0;
}

int efd_send(int sock, uint32_t type, const void *payload, size_t len, const int *fds, int nfds) {
  efd_hdr h = {EFD_MAGIC, type, len};
  if (nfds > 0) {                               /* the header travels with the descriptors as ancillary data */
    struct iovec iov = {&h, sizeof h};
    char ctl[CMSG_SPACE(sizeof(int) * PCE_MAX_SEGMENTS)];
    memset(ctl, 0, sizeof ctl);
    struct msghdr m = {0};
    m.msg_iov = &iov; m.msg_iovlen = 1; m.msg_control = ctl; m.msg_controllen = CMSG_SPACE(sizeof(int) * (size_t)nfds);
    struct cmsghdr *c = CMSG_FIRSTHDR(&m);
    c->cmsg_level = SOL_SOCKET; c->cmsg_type = SCM_RIGHTS; c->cmsg_len = CMSG_LEN(sizeof(int) * (size_t)nfds);
    memcpy(CMSG_DATA(c), fds, sizeof(int) * (size_t)nfds);
    ssize_t k;
    do k = sendmsg(sock, &m, MSG_NOSIGNAL); while (k < 0 && errno == EINTR);
    if (k != (ssize_t)sizeof h) return -1;
  } else if (write_all(sock, &h, sizeof h)) return -1;
  return len ? write_all(sock, payload, len) : 0;
}

int efd_recv(int sock, uint32_t *type, void **payload, size_t *len, int *fds, int max_fds, int *nfds) {
  efd_hdr h;
  struct iovec iov = {&h, sizeof h};
  char ctl[CMSG_SPACE(sizeof(int) * PCE_MAX_SEGMENTS)];
  struct msghdr m = {0};
  m.msg_iov = &iov; m.msg_iovlen = 1; m.msg_control = ctl; m.msg_controllen = sizeof ctl;
  ssize_t k;
  do k = recvmsg(sock, &m, MSG_CMSG_CLOEXEC); while (k < 0 && errno == EINTR);
  if (k <= 0) return -1;
  if ((size_t)k < sizeof h && read_all(sock, (char *)&h + k, sizeof h - (size_t)k)) return -1;
  if (nfds) *nfds = 0;
  for (struct cmsghdr *c = CMSG_FIRSTHDR(&m); c; c = CMSG_NXTHDR(&m, c))
    if (c->cmsg_level == SOL_SOCKET && c->cmsg_type == SCM_RIGHTS) {
      const int n = (int)((c->cmsg_len - CMSG_LEN(0)) / sizeof(int));
      const int *src = (const int *)CMSG_DATA(c);
      for (int i = 0; i < n; ++i) {
        if (fds && nfds && *nfds < max_fds) fds[(*nfds)++] = src[i]; else close(src[i]);
      }
    }
  if (h.magic != EFD_MAGIC || h.len > ((uint64_t)1 << 32)) return -1;
  *type = h.type; *len = (size_t)h.len; *payload = NULL;
  if (h.len) {
    *payload = malloc((size_t)h.len);
    if (!*payload || read_all(sock, *payload, (size_t)h.len)) { free(*payload); *payload = NULL; return -1; }
  }
  return 0;
}

const char *efd_default_socket(char *buf, size_t n) {
  const char *e = getenv("EST_FACTD_SOCKET");
  if (e && *e) snprintf(buf, n, "%s", e);
  else if (strncmp(pc_engine_backend(), "cuda", 4) == 0) snprintf(buf, n, "/tmp/est-factd-%u.sock", (unsigned)getuid());
  else snprintf(buf, n, "/tmp/est-factd-%s-%u.sock", pc_engine_backend(), (unsigned)getuid());
  return buf;
}

/* ---- in-process engine (one per process) -------------------------------------------------------------------------- */
static pc_engine *g_engine;
static pthread_mutex_t g_engine_mu = PTHREAD_MUTEX_INITIALIZER;

static ef_conn *open_inproc(const int *devices, int ndev, const ef_conn_req *req, char *err, size_t errlen) {
  pthread_mutex_lock(&g_engine_mu);
  if (!g_engine) {
    int dflt = 0;
    /* the first segment is sized for this job alone: pinning memory is the slow part of start-up */
    const size_t slab = (size_t)req->arena_cap + (size_t)req->var_cap + (sizeof(pc_job) + 4u * PC_RES_INTS) * (size_t)req->jobs_cap + (64u << 10);
    g_engine = pc_engine_create(ndev > 0 ? devices : &dflt, ndev > 0 ? ndev : 1, PCE_HDR_BYTES + slab * (size_t)req->nlanes + (1u << 20));
  }
  pc_engine *e = g_engine;
  pthread_mutex_unlock(&g_engine_mu);
  if (!e) { snprintf(err, errlen, "%s", pc_last_error()); return NULL; }
  pc_session_req r = {req->device < 0 || ndev <= 0 ? 0 : req->device, req->genome, req->genome_len, req->word_len, req->depth_rate, req->nlanes,
                      req->arena_cap, req->var_cap, req->jobs_cap};
  pc_session_info info;
  if (pc_engine_open(e, &r, &info)) { snprintf(err, errlen, "%s", pc_last_error()); return NULL; }
  if (req->timers) pc_engine_enable_timers(e, 1);
  ef_conn *c = calloc(1, sizeof *c);
  c->daemon = false; c->eng = e; c->sock = -1; c->session = info.session; c->gpu = info.gpu; c->nlanes = info.nlanes;
  memcpy(c->lane, info.lane, sizeof(uint32_t) * (size_t)info.nlanes);
  c->hdr = pc_engine_segment_base(e, info.gpu, 0);
  pthread_mutex_init(&c->mu, NULL);
  return c;
}

/* ---- the server: connect, or start one ---------------------------------------------------------------------------- */
static int connect_path(const char *path) {
  int s = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
  if (s < 0) return -1;
  struct sockaddr_un a;
  memset(&a, 0, sizeof a);
  a.sun_family = AF_UNIX;
  { const size_t n = strlen(path); memcpy(a.sun_path, path, n < sizeof a.sun_path - 1 ? n : sizeof a.sun_path - 1); }
  if (connect(s, (struct sockaddr *)&a, sizeof a) != 0) { close(s); return -1; }
  return s;
}

static double mono(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec; }

/* est-factd lives next to this executable.  Double fork: the server must not be a child of the job (pintron.py waits
 * for est-fact's process group) and must not hold its stdio. */
static int spawn_server(const char *path) {
  char exe[4096];
  ssize_t n = readlink("/proc/self/exe", exe, sizeof exe - 16);
  if (n <= 0) return -1;
  exe[n] = 0;
  char *slash = strrchr(exe, '/');
  if (!slash) return -1;
  strcpy(slash + 1, "est-factd");
  if (access(exe, X_OK) != 0) return -1;
  pid_t p = fork();
  if (p < 0) return -1;
  if (p == 0) {
    setsid();
    if (fork() != 0) _exit(0);
    char log[256];
    snprintf(log, sizeof log, "/tmp/est-factd-%u.log", (unsigned)getuid());
    int fd = open(log, O_WRONLY | O_CREAT | O_APPEND, 0600), nul = open("/dev/null", O_RDONLY);
    if (nul >= 0) dup2(nul, 0);
    if (fd >= 0) { dup2(fd, 1); dup2(fd, 2); }
    for (int k = 3; k < 256; ++k) close(k);
    if (chdir("/") != 0) _exit(1);
    const char *idle = getenv("EST_FACTD_IDLE");
    execl(exe, "est-factd", "--socket", path, "--idle-timeout", idle && *idle ? idle : "300", (char *)NULL);
    _exit(127);
  }
  int st = 0;
  waitpid(p, &st, 0);
  return 0;
}

static ef_conn *open_daemon(bool may_spawn, int ordinal, const ef_conn_req *req, char *err, size_t errlen) {
  char path[256];
  efd_default_socket(path, sizeof path);
  int s = connect_path(path);
  if (s < 0 && may_spawn && !getenv("EST_FACT_NO_SPAWN")) {
    if (spawn_server(path) == 0) {
      const double t0 = mono();
      while (s < 0 && mono() - t0 < 90.0) {       /* a cold CUDA context on a fresh box can take several seconds */
        struct timespec ts = {0, 20 * 1000 * 1000};
        nanosleep(&ts, NULL);
        s = connect_path(path);
      }
    }
  }
  if (s < 0) { snprintf(err, errlen, "no est-factd answering on %s", path); return NULL; }
  efd_hello h = {ordinal, req->word_len, req->nlanes, req->timers ? 1 : 0, req->jobs_cap, 0, req->arena_cap, req->var_cap,
                 (uint64_t)req->genome_len, req->depth_rate};
  const size_t len = sizeof h + req->genome_len;
  char *buf = malloc(len);
  memcpy(buf, &h, sizeof h);
  memcpy(buf + sizeof h, req->genome, req->genome_len);
  int rc = efd_send(s, EFD_HELLO, buf, len, NULL, 0);
  free(buf);
  uint32_t type = 0; void *pl = NULL; size_t plen = 0; int fds[PCE_MAX_SEGMENTS], nfds = 0;
  if (rc || efd_recv(s, &type, &pl, &plen, fds, PCE_MAX_SEGMENTS, &nfds)) { snprintf(err, errlen, "est-factd closed the connection during the handshake"); close(s); return NULL; }
  if (type != EFD_HELLO_OK || plen != sizeof(efd_hello_ok)) {
    snprintf(err, errlen, "est-factd: %.*s", (int)(type == EFD_ERROR ? plen : 16), type == EFD_ERROR ? (char *)pl : "unexpected reply");
    free(pl); close(s);
    for (int i = 0; i < nfds; ++i) close(fds[i]);
    return NULL;
  }
  efd_hello_ok *ok = pl;
  ok->backend[sizeof ok->backend - 1] = 0;
  if (strcmp(ok->backend, pc_engine_backend()) != 0) {
    snprintf(err, errlen, "the est-factd on %s is a '%s' server, this est-fact wants '%s'", path, ok->backend, pc_engine_backend());
    free(pl); close(s);
    for (int i = 0; i < nfds; ++i) close(fds[i]);
    return NULL;
  }
  ef_conn *c = calloc(1, sizeof *c);
  c->daemon = true; c->sock = s; c->session = ok->session; c->gpu = ok->gpu; c->nlanes = ok->nlanes; c->nsegs = ok->nsegs;
  memcpy(c->lane, ok->lane, sizeof(uint32_t) * (size_t)ok->nlanes);
  for (int i = 0; i < PCE_MAX_SEGMENTS; ++i) c->seg_fd[i] = -1;
  for (int i = 0; i < ok->nsegs && i < nfds; ++i) { c->seg_fd[i] = fds[i]; c->seg_bytes[i] = (size_t)ok->seg_bytes[i]; }
  free(pl);
  pthread_mutex_init(&c->mu, NULL);
  c->hdr = (pce_hdr *)efc_seg(c, 0);
  if (!c->hdr || c->hdr->magic != PCE_MAGIC || c->hdr->version != PCE_VERSION) {
    snprintf(err, errlen, "est-factd speaks another lane protocol version (or the lane table could not be mapped)");
    close(s); free(c);
    return NULL;
  }
  return c;
}

ef_conn *efc_open(const char *mode, const int *devices, int ndev, const ef_conn_req *req, char *err, size_t errlen) {
  err[0] = 0;
  const int *in_devices = devices;                    /* in-process ordinals: 0..n-1 when main() narrowed CUDA_VISIBLE_DEVICES to the list */
  int seq[16];
  if (ndev > 0 && getenv("EF_DEVICES_NARROWED")) { for (int i = 0; i < ndev && i < 16; ++i) seq[i] = i; in_devices = seq; }
  if (mode && strcmp(mode, "inproc") == 0) return open_inproc(in_devices, ndev, req, err, errlen);
  const bool strict = mode && strcmp(mode, "daemon") == 0;
  /* req->device is a position in `devices`; the server wants the ordinal itself (-1: its least loaded GPU) */
  ef_conn *c = open_daemon(true, (ndev > 0 && req->device >= 0 && req->device < ndev) ? devices[req->device] : -1, req, err, errlen);
  if (c || strict) return c;
  char err2[256];
  c = open_inproc(in_devices, ndev, req, err2, sizeof err2);
  if (!c) { const size_t l = strlen(err); snprintf(err + l, errlen - l, "; in-process engine: %s", err2); }
  return c;
}

uint8_t *efc_seg(ef_conn *c, uint32_t seg) {
  if (seg >= PCE_MAX_SEGMENTS) return NULL;
  if (!c->daemon) return pc_engine_segment_base(c->eng, c->gpu, (int)seg);
  if (c->seg_base[seg]) return c->seg_base[seg];
  if (c->seg_fd[seg] < 0) return NULL;
  void *p = mmap(NULL, c->seg_bytes[seg], PROT_READ | PROT_WRITE, MAP_SHARED, c->seg_fd[seg], 0);
  if (p == MAP_FAILED) return NULL;
  c->seg_base[seg] = p;
  return p;
}

int efc_resize(ef_conn *c, int k, uint64_t arena_cap, uint32_t jobs_cap, uint64_t var_cap, uint64_t keep_arena, uint32_t keep_jobs) {
  if (!c->daemon) return pc_engine_resize_lane(c->eng, c->session, c->lane[k], arena_cap, jobs_cap, var_cap, keep_arena, keep_jobs);
  pthread_mutex_lock(&c->mu);
  efd_resize r = {c->lane[k], jobs_cap, keep_jobs, (uint32_t)c->nsegs, arena_cap, var_cap, keep_arena};
  uint32_t type = 0; void *pl = NULL; size_t plen = 0; int fds[PCE_MAX_SEGMENTS], nfds = 0;
  int rc = efd_send(c->sock, EFD_RESIZE, &r, sizeof r, NULL, 0) || efd_recv(c->sock, &type, &pl, &plen, fds, PCE_MAX_SEGMENTS, &nfds);
  if (!rc && type == EFD_RESIZE_OK && plen == sizeof(efd_resize_ok)) {
    efd_resize_ok *ok = pl;
    for (int i = 0; i < nfds && c->nsegs + i < ok->nsegs; ++i) { c->seg_fd[c->nsegs + i] = fds[i]; c->seg_bytes[c->nsegs + i] = (size_t)ok->seg_bytes[c->nsegs + i]; }
    c->nsegs = ok->nsegs;
  } else {
    if (!rc && type == EFD_ERROR) fprintf(stderr, "* ERROR est-factd: %.*s\n", (int)plen, (char *)pl);
    rc = -1;
  }
  free(pl);
  pthread_mutex_unlock(&c->mu);
  return rc;
}

int efc_alive(void *conn) {
  ef_conn *c = conn;
  if (!c->daemon) return 1;
  struct pollfd p = {c->sock, POLLIN, 0};
  if (poll(&p, 1, 0) > 0 && (p.revents & (POLLHUP | POLLERR | POLLNVAL))) return 0;
  if (p.revents & POLLIN) { char b; if (recv(c->sock, &b, 1, MSG_PEEK | MSG_DONTWAIT) == 0) return 0; }
  return 1;
}

int efc_close(ef_conn *c, pc_session_stats *stats) {
  if (!c) return 0;
  int rc = 0;
  if (stats) memset(stats, 0, sizeof *stats);
  if (!c->daemon) rc = pc_engine_close(c->eng, c->session, stats);
  else {
    uint32_t type = 0; void *pl = NULL; size_t plen = 0;
    if (efd_send(c->sock, EFD_BYE, NULL, 0, NULL, 0) == 0 && efd_recv(c->sock, &type, &pl, &plen, NULL, 0, NULL) == 0 && type == EFD_STATS &&
        plen == sizeof(pc_session_stats)) { if (stats) memcpy(stats, pl, sizeof *stats); }
    else rc = -1;
    free(pl);
    close(c->sock);
    /* the mappings go with the process */
  }
  free(c);
  return rc;
}
