/* ref_text.c -- TEST INFRASTRUCTURE (part of liboracle_sts.so): fast access to the text files the
 * unmodified reference driver writes with `--output 2`.
 *
 * The reference has one output channel for a state: UserOutput::write prints every value of the local
 * block with 16 significant digits ("%.15e ", diffusion_2D/diffusion_2D.cpp:760-761, :819-824), one
 * line per output time, into diffusion_2d_solution.<rank>.txt (header lines start with '#').  At the
 * kernel geometries the bench runs (16384 cells wide) a state is 10^7 .. 10^8 values: parsing and
 * re-printing that in Python takes minutes, here it takes seconds.  Nothing in the product loads this. */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <pthread.h>
#include <string.h>
#include <unistd.h>

/* Parse the LAST data line of a solution file: "t v0 v1 ... v(n-1)".  Returns the number of values
 * (without t) written to out (at most nmax), or -1 on error; *t receives the time stamp. */
long orc_text_last_line(const char* path, double* out, long nmax, double* t)
{
  FILE* f = fopen(path, "rb");
  if (!f) return -1;
  if (fseek(f, 0, SEEK_END) != 0) { fclose(f); return -1; }
  const long len = ftell(f);
  if (len <= 0) { fclose(f); return -1; }
  /* find the start of the last non-empty line by scanning backwards in blocks */
  enum { BLK = 1 << 16 };
  char blk[BLK];
  long end = len, beg = -1;
  int seen_text = 0;
  for (long pos = len; pos > 0 && beg < 0;)
  {
    const long lo = pos > BLK ? pos - BLK : 0;
    if (fseek(f, lo, SEEK_SET) != 0 || fread(blk, 1, (size_t)(pos - lo), f) != (size_t)(pos - lo)) { fclose(f); return -1; }
    for (long k = pos - 1; k >= lo; k--)
    {
      const char c = blk[k - lo];
      if (!seen_text)
      {
        if (c == '\n' || c == ' ' || c == '\r') { end = k; continue; }
        seen_text = 1;
      }
      else if (c == '\n') { beg = k + 1; break; }
    }
    pos = lo;
  }
  if (!seen_text) { fclose(f); return -1; }
  if (beg < 0) beg = 0;
  char* line = (char*)malloc((size_t)(end - beg) + 1);
  if (!line) { fclose(f); return -1; }
  if (fseek(f, beg, SEEK_SET) != 0 || fread(line, 1, (size_t)(end - beg), f) != (size_t)(end - beg)) { free(line); fclose(f); return -1; }
  fclose(f);
  line[end - beg] = '\0';
  long n = -1;
  if (line[0] != '#')
  {
    char* p = line;
    char* q = NULL;
    *t      = strtod(p, &q);
    n       = 0;
    if (q != p)
    {
      p = q;
      while (n < nmax)
      {
        const double v = strtod(p, &q);
        if (q == p) break;
        out[n++] = v;
        p        = q;
      }
    }
  }
  free(line);
  return n;
}

/* How many of ours[i] do NOT print ("%.15e") to the text the reference printed for ref[i]?  ref[] holds the
 * reference's printed values parsed back (orc_text_last_line), so printing ref[i] again reproduces the
 * reference's characters: two doubles "equal the reference's output" iff their 16-digit strings agree. */
typedef struct { const double* ours; const double* ref; long lo, hi, bad; } mm_job;
static void* mm_worker(void* arg)
{
  mm_job* j = (mm_job*)arg;
  long bad  = 0;
  for (long i = j->lo; i < j->hi; i++)
  {
    if (j->ours[i] == j->ref[i]) continue; /* the printed value happens to be exactly ours */
    char a[40], b[40];
    snprintf(a, sizeof(a), "%.15e", j->ours[i]);
    snprintf(b, sizeof(b), "%.15e", j->ref[i]);
    if (strcmp(a, b) != 0) bad++;
  }
  j->bad = bad;
  return NULL;
}
long orc_text_print_mismatches(const double* ours, const double* ref, long n)
{
  enum { MAXT = 64 };
  long nt = sysconf(_SC_NPROCESSORS_ONLN);
  if (nt < 1) nt = 1;
  if (nt > MAXT) nt = MAXT;
  if (n < 100000) nt = 1;
  pthread_t th[MAXT];
  mm_job job[MAXT];
  for (long k = 0; k < nt; k++)
  {
    job[k].ours = ours; job[k].ref = ref; job[k].bad = 0;
    job[k].lo = n * k / nt; job[k].hi = n * (k + 1) / nt;
    if (pthread_create(&th[k], NULL, mm_worker, &job[k]) != 0) { mm_worker(&job[k]); th[k] = 0; }
  }
  long bad = 0;
  for (long k = 0; k < nt; k++)
  {
    if (th[k]) pthread_join(th[k], NULL);
    bad += job[k].bad;
  }
  return bad;
}
