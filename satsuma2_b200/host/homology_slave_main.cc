// HomologyByXCorrSlave on B200: drop-in for the reference's slave process
// (analysis/HomologyByXCorrSlave.cc:331-533).  Same flags, same chunking (target overlap
// t_chunk/4, query overlap 0), same TCP exchange with the master's WorkQueue
// (analysis/WorkQueue.cc:71-166), byte-identical t_pair / t_result structs:
//   per connection:  -> uint32 slave_id, uint32 n, n x t_result(72 B)
//                    <- int32 count (-1 = terminate), count x t_pair(28 B)
// Differences that are deliberate: reads and writes are looped until complete (the reference
// issues single read()/write() calls), all received blocks are aligned in one batched GPU call
// instead of `-p` worker threads, and an empty answer is retried after 200 ms instead of 10 s.
#include <arpa/inet.h>
#include <netdb.h>
#include <netinet/in.h>
#include <sys/socket.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>

#include "sx_host.h"

using namespace sxh;

static bool write_all(int fd, const void *p, size_t n) {
  const char *c = (const char *)p;
  while (n) {
    ssize_t w = write(fd, c, n);
    if (w <= 0) return false;
    c += w;
    n -= (size_t)w;
  }
  return true;
}
static bool read_all(int fd, void *p, size_t n) {
  char *c = (char *)p;
  while (n) {
    ssize_t r = read(fd, c, n);
    if (r <= 0) return false;
    c += r;
    n -= (size_t)r;
  }
  return true;
}

static const char *flag(std::map<std::string, std::string> &a, const char *k, const char *def) {
  auto it = a.find(k);
  return it == a.end() ? def : it->second.c_str();
}
static bool truthy(const char *v) { return !(strcmp(v, "0") == 0 || strcmp(v, "false") == 0); }

int main(int argc, char **argv) {
  std::map<std::string, std::string> a;
  for (int i = 1; i + 1 < argc; i += 2) a[argv[i]] = argv[i + 1];
  const std::string master = flag(a, "-master", ""), q = flag(a, "-q", ""), t = flag(a, "-t", "");
  if (master.empty() || q.empty() || t.empty()) {
    fprintf(stderr,
            "usage: %s -master <host> [-port 3491] -sid <id> -q <query fasta> -t <target fasta> [-l 0]\n"
            "          [-q_chunk 4096] [-t_chunk 4096] [-min_prob 0.9999] [-cutoff 1.8] [-cutoff_fast 2.9]\n"
            "          [-prob_table 0] [-debug_targets 0] [-device 0]   (-p is accepted and ignored)\n",
            argv[0]);
    return 2;
  }
  const int port = atoi(flag(a, "-port", "3491"));
  const unsigned int slave_id = (unsigned int)atoi(flag(a, "-sid", "0"));
  const unsigned int debug_max = (unsigned int)atoi(flag(a, "-debug_targets", "0"));
  HomologyByXCorr::Options opt;
  opt.device = atoi(flag(a, "-device", "0"));
  opt.t_chunk = atoi(flag(a, "-t_chunk", "4096"));
  opt.q_chunk = atoi(flag(a, "-q_chunk", "4096"));
  opt.cutoff = atof(flag(a, "-cutoff", "1.8"));
  opt.cutoff_fast = atof(flag(a, "-cutoff_fast", "2.9"));
  opt.min_len = atoi(flag(a, "-l", "0"));
  opt.min_prob_flag = atof(flag(a, "-min_prob", "0.9999"));
  opt.prob_table = truthy(flag(a, "-prob_table", "0"));

  std::vector<Sequence> qs, ts;
  std::string err;
  printf("Loading query sequence:  %s\n", q.c_str());
  if (!read_fasta(q, qs, &err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
  printf("Loading target sequence:  %s\n", t.c_str());
  if (!read_fasta(t, ts, &err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
  ChunkList tc, qc;
  chunk_sequences(qs, opt.q_chunk, 0, 0, 0, qc);               // Slave.cc:391-392
  chunk_sequences(ts, opt.t_chunk, opt.t_chunk / 4, 0, 0, tc);  // Slave.cc:400-401
  HomologyByXCorr hx;
  if (!hx.init(opt, tc, qc)) { fprintf(stderr, "slave: %s\n", hx.error().c_str()); return 1; }
  printf("chunks: target %d query %d, target total %.0f\n== Entering communication loop ==\n", tc.n(), qc.n(),
         hx.target_total());

  struct hostent *server = gethostbyname(master.c_str());
  if (!server) { fprintf(stderr, "ERROR resolving master hostname\n"); return 1; }
  struct sockaddr_in addr;
  memset(&addr, 0, sizeof(addr));
  addr.sin_family = AF_INET;
  memcpy(&addr.sin_addr.s_addr, server->h_addr, (size_t)server->h_length);
  addr.sin_port = htons((uint16_t)port);

  std::vector<t_result> results;
  std::vector<t_pair> pairs;
  unsigned int total_targets = 0;
  bool finished = false;
  while (!finished) {
    int fd = socket(AF_INET, SOCK_STREAM, 0);
    if (fd < 0 || connect(fd, (struct sockaddr *)&addr, sizeof(addr)) < 0) {
      if (fd >= 0) close(fd);
      sleep(1);
      continue;
    }
    const unsigned int n = (unsigned int)results.size();
    int count = 0;
    bool ok = write_all(fd, &slave_id, sizeof(slave_id)) && write_all(fd, &n, sizeof(n)) &&
              (n == 0 || write_all(fd, results.data(), sizeof(t_result) * n)) && read_all(fd, &count, sizeof(count));
    if (ok) results.clear();
    pairs.clear();
    if (ok && count > 0) {
      pairs.resize((size_t)count);
      ok = read_all(fd, pairs.data(), sizeof(t_pair) * (size_t)count);
    }
    close(fd);
    if (!ok) { sleep(1); continue; }
    if (count == -1) finished = true;
    if (count > 0) {
      if (!hx.align_targets(pairs.data(), count, results)) {
        fprintf(stderr, "slave: %s\n", hx.error().c_str());
        return 1;
      }
      total_targets += (unsigned int)count;
      if (debug_max && total_targets >= debug_max) finished = true;
    } else if (count == 0) {
      usleep(200000);
    }
  }
  if (!results.empty()) {  // -debug_targets stop: hand the last records over before leaving
    int fd = socket(AF_INET, SOCK_STREAM, 0);
    if (fd >= 0 && connect(fd, (struct sockaddr *)&addr, sizeof(addr)) == 0) {
      const unsigned int n = (unsigned int)results.size();
      int count = 0;
      if (write_all(fd, &slave_id, sizeof(slave_id)) && write_all(fd, &n, sizeof(n)) &&
          write_all(fd, results.data(), sizeof(t_result) * n) && read_all(fd, &count, sizeof(count)) && count > 0) {
        pairs.resize((size_t)count);
        read_all(fd, pairs.data(), sizeof(t_pair) * (size_t)count);  // drained and dropped
      }
    }
    if (fd >= 0) close(fd);
  }
  printf("== Processing finished ==\n");
  return 0;
}
