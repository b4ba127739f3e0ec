/* engine_client.h — how the est-fact host reaches the batch engine (include/pintron_engine.h): either an engine inside
 * this process, or the resident server est-factd over a UNIX socket + shared-memory lanes.  Also the wire format the
 * two programs share.  Plain C; no CUDA on this side of the socket. */
#ifndef EF_ENGINE_CLIENT_H
#define EF_ENGINE_CLIENT_H
#include <pthread.h>
#include <stdbool.h>
#include "pintron_engine.h"

/* ---- wire format (est-fact <-> est-factd), little-endian, one header + payload per message ---------------------- */
#define EFD_MAGIC 0x45464432u   /* "EFD2" */
enum { EFD_HELLO = 1, EFD_HELLO_OK = 2, EFD_ERROR = 3, EFD_RESIZE = 4, EFD_RESIZE_OK = 5, EFD_BYE = 6, EFD_STATS = 7, EFD_SHUTDOWN = 8 };
typedef struct efd_hdr { uint32_t magic, type; uint64_t len; } efd_hdr;
typedef struct efd_hello {
  int32_t gpu, word_len, nlanes, timers;
  uint32_t jobs_cap, pad;
  uint64_t arena_cap, var_cap, genome_len;
  double depth_rate;
} efd_hello;                                     /* followed by genome_len bytes */
typedef struct efd_hello_ok {
  uint32_t session; int32_t gpu, nlanes, nsegs;  /* nsegs file descriptors ride along (SCM_RIGHTS) */
  char backend[16];                              /* pc_engine_backend() of the server */
  uint64_t seg_bytes[PCE_MAX_SEGMENTS];
  uint32_t lane[PCE_MAX_SESSION_LANES];
} efd_hello_ok;
typedef struct efd_resize { uint32_t lane, jobs_cap, keep_jobs, have_segs; uint64_t arena_cap, var_cap, keep_arena; } efd_resize;
typedef struct efd_resize_ok { int32_t nsegs, new_fds; uint64_t seg_bytes[PCE_MAX_SEGMENTS]; } efd_resize_ok;   /* fds of segments have_segs .. nsegs-1 ride along */

int efd_send(int sock, uint32_t type, const void *payload, size_t len, const int *fds, int nfds);
/* receives one message; *payload is malloc'ed (caller frees); fds (up to max_fds) are stored, *nfds set */
int efd_recv(int sock, uint32_t *type, void **payload, size_t *len, int *fds, int max_fds, int *nfds);
const char *efd_default_socket(char *buf, size_t n);      /* $EST_FACTD_SOCKET or /tmp/est-factd-<uid>.sock (test stand-in: est-factd-<backend>-<uid>.sock) */

/* ---- the client handle ------------------------------------------------------------------------------------------- */
typedef struct ef_conn {
  bool daemon;                      /* false: engine inside this process */
  pc_engine *eng;
  int sock;
  uint32_t session;
  int gpu, nlanes;
  uint32_t lane[PCE_MAX_SESSION_LANES];
  pce_hdr *hdr;
  uint8_t *seg_base[PCE_MAX_SEGMENTS];
  size_t seg_bytes[PCE_MAX_SEGMENTS];
  int seg_fd[PCE_MAX_SEGMENTS];
  int nsegs;
  pthread_mutex_t mu;               /* resize requests from several worker threads */
} ef_conn;

typedef struct ef_conn_req {
  int device;                       /* position in the device list given to efc_open (-1 / empty list: any GPU) */
  const char *genome; size_t genome_len; int word_len; double depth_rate;
  int nlanes; uint64_t arena_cap, var_cap; uint32_t jobs_cap;
  bool timers;
} ef_conn_req;

/* mode: "inproc", "daemon" (fail when no server answers and none can be started) or "auto" (server if reachable or
 * startable, else in-process).  devices / ndev: CUDA ordinals for an in-process engine (shared by the connections of
 * this process).  On failure returns NULL with a message in err. */
ef_conn *efc_open(const char *mode, const int *devices, int ndev, const ef_conn_req *req, char *err, size_t errlen);
static inline pce_lane *efc_lane(ef_conn *c, int k) { return &c->hdr->lanes[c->lane[k]]; }
uint8_t *efc_seg(ef_conn *c, uint32_t seg);                /* maps the segment on first use (daemon mode) */
int efc_resize(ef_conn *c, int k, uint64_t arena_cap, uint32_t jobs_cap, uint64_t var_cap, uint64_t keep_arena, uint32_t keep_jobs);
int efc_alive(void *conn);                                 /* for pce_wait: false once the server has gone away */
int efc_close(ef_conn *c, pc_session_stats *stats);

#endif
