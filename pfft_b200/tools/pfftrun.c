/* pfftrun -- start N ranks of a minimpi program on this node (the `mpirun -np N` of
 * pfft_b200; the reference's tests are launched with mpirun, tests/run_checks.sh:80-101).
 *
 *   pfftrun -np 4 ./simple_check_c2c [args...]
 *
 * Each rank gets PFFT_MPI_JOB / PFFT_MPI_RANK / PFFT_MPI_SIZE (read by MPI_Init in
 * libpfft_b200.so) and PFFT_B200_AUTODEVICE=1 (rank r uses GPU r % device_count, so
 * N ranks also run on a single GPU).  If any rank fails the others are terminated and
 * the launcher exits non-zero; a wall-clock limit (-timeout S, default 600 s) guards
 * against hangs.
 */
#define _GNU_SOURCE
#include <errno.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

static pid_t *kids;
static int nkids;
static char shm_name[128];

static void kill_all(int sig) {
  for (int i = 0; i < nkids; i++)
    if (kids[i] > 0) kill(kids[i], sig);
}

static void on_alarm(int s) {
  (void)s;
  fprintf(stderr, "pfftrun: time limit reached, terminating ranks\n");
  kill_all(SIGKILL);
  shm_unlink(shm_name);
  _exit(124);
}

static void on_term(int s) {
  kill_all(SIGTERM);
  shm_unlink(shm_name);
  _exit(128 + s);
}

int main(int argc, char **argv) {
  int np = 1, limit = 600, a = 1;
  while (a < argc && argv[a][0] == '-') {
    if ((!strcmp(argv[a], "-np") || !strcmp(argv[a], "-n")) && a + 1 < argc) { np = atoi(argv[a + 1]); a += 2; }
    else if (!strcmp(argv[a], "-timeout") && a + 1 < argc) { limit = atoi(argv[a + 1]); a += 2; }
    else if (!strcmp(argv[a], "--")) { a++; break; }
    else break;
  }
  if (a >= argc || np < 1 || np > 64) {
    fprintf(stderr, "usage: pfftrun -np N [-timeout S] program [args...]\n");
    return 2;
  }
  char job[64];
  snprintf(job, sizeof job, "%d_%ld", (int)getpid(), (long)time(NULL));
  snprintf(shm_name, sizeof shm_name, "/pfftb200_%s", job);
  kids = calloc((size_t)np, sizeof(pid_t));
  nkids = np;
  signal(SIGALRM, on_alarm);
  signal(SIGTERM, on_term);
  signal(SIGINT, on_term);
  for (int r = 0; r < np; r++) {
    pid_t p = fork();
    if (p < 0) { perror("fork"); kill_all(SIGKILL); return 1; }
    if (p == 0) {
      char buf[32];
      setenv("PFFT_MPI_JOB", job, 1);
      snprintf(buf, sizeof buf, "%d", r);
      setenv("PFFT_MPI_RANK", buf, 1);
      snprintf(buf, sizeof buf, "%d", np);
      setenv("PFFT_MPI_SIZE", buf, 1);
      if (!getenv("PFFT_B200_DEVICE")) setenv("PFFT_B200_AUTODEVICE", "1", 1);
      execvp(argv[a], argv + a);
      fprintf(stderr, "pfftrun: cannot exec %s: %s\n", argv[a], strerror(errno));
      _exit(127);
    }
    kids[r] = p;
  }
  alarm((unsigned)limit);
  int rc = 0, left = np;
  while (left > 0) {
    int st = 0;
    pid_t p = wait(&st);
    if (p < 0) break;
    for (int i = 0; i < np; i++)
      if (kids[i] == p) kids[i] = 0;
    left--;
    int code = WIFEXITED(st) ? WEXITSTATUS(st) : 128 + WTERMSIG(st);
    if (code != 0 && rc == 0) {
      rc = code;
      kill_all(SIGTERM);
    }
  }
  shm_unlink(shm_name);
  return rc;
}
