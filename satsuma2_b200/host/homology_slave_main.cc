// HomologyByXCorrSlave on B200: drop-in for the reference's slave process
// (analysis/HomologyByXCorrSlave.cc:331-533).  Same flags, same chunking (target overlap
// t_chunk/4, query overlap 0), same TCP exchange with the master's WorkQueue
// (analysis/WorkQueue.cc:71-166), byte-identical t_pair / t_result structs:
//   per connection:  -> uint32 slave_id, uint32 n, n x t_result(72 B)
//                    <- int32 count (-1 = terminate), count x t_pair(28 B)
// Structure: the reference runs `-p` CPU workers that pop one t_pair at a time while the main thread exchanges with
// the master whenever fewer than 2*p pairs are queued (Slave.cc:447-526).  Here ONE GPU worker thread takes
// everything that is queued and aligns it in one batched call, while the main thread keeps exchanging: it ships the
// records of the previous batch and keeps asking for more pairs until `-queue` blocks (default 64) are waiting -- the
// master hands out only 2 x its per-slave thread count per connection (WorkQueue.cc:100-107), so a GPU slave fills its
// queue over several back-to-back connections.  Exchanges and GPU work overlap; nothing the master hands out is
// dropped (pairs received in the last exchanges are still aligned and their records delivered).
// Other deliberate differences: reads and writes are looped until complete (the reference issues single
// read()/write() calls), an empty answer is retried after 50 ms instead of 10 s, `-device` (default: (sid-1) mod
// number of GPUs, so the unmodified master's `-sid 1..n` slaves land on different GPUs of a box) and `-gpus n`
// (this one process drives n GPUs through sx_multi, target list split by range).
#include <arpa/inet.h>
#include <netdb.h>
#include <netinet/in.h>
#include <sys/socket.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

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
            "          [-prob_table 0] [-debug_targets 0] [-device d | -gpus n] [-queue 64]   (-p is accepted and ignored)\n",
            argv[0]);
    return 2;
  }
  const int port = atoi(flag(a, "-port", "3491"));
  const unsigned int slave_id = (unsigned int)atoi(flag(a, "-sid", "0"));
  const unsigned int debug_max = (unsigned int)atoi(flag(a, "-debug_targets", "0"));
  const size_t queue_target = (size_t)std::max(1, atoi(flag(a, "-queue", "64")));
  HomologyByXCorr::Options opt;
  const int ndev = sx_device_count();
  opt.device = a.count("-device") ? atoi(a["-device"].c_str()) : (ndev > 0 ? (int)((slave_id > 0 ? slave_id - 1 : 0) % (unsigned)ndev) : 0);
  opt.n_gpus = atoi(flag(a, "-gpus", "0"));
  opt.t_chunk = atoi(flag(a, "-t_chunk", "4096"));
  opt.q_chunk = atoi(flag(a, "-q_chunk", "4096"));
  opt.cutoff = atof(flag(a, "-cutoff", "1.8"));
  opt.cutoff_fast = atof(flag(a, "-cutoff_fast", "2.9"));
  opt.min_len = atoi(flag(a, "-l", "0"));
  opt.min_prob_flag = atof(flag(a, "-min_prob", "0.9999"));
  opt.prob_table = truthy(flag(a, "-prob_table", "0"));
  if (opt.t_chunk < 1 || opt.q_chunk < 1) {  // before any chunk arithmetic divides by them
    fprintf(stderr, "slave: -t_chunk / -q_chunk must be positive\n");
    return 2;
  }

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

  // shared between the exchange loop (this thread) and the GPU worker
  std::mutex mu;
  std::condition_variable cv;
  std::deque<t_pair> queue;        // pairs received, not yet taken by the worker
  std::vector<t_result> results;   // records computed, not yet delivered
  size_t in_flight = 0;            // pairs the worker is aligning right now
  bool stop = false, failed = false;
  unsigned long long exchanges = 0, batches = 0, blocks_done = 0;

  std::thread worker([&]() {
    std::vector<t_pair> batch;
    std::vector<t_result> out;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return stop || !queue.empty(); });
        if (queue.empty()) return;  // stop and nothing left to do
        batch.assign(queue.begin(), queue.end());
        queue.clear();
        in_flight = batch.size();
      }
      cv.notify_all();
      out.clear();
      const bool ok = hx.align_targets(batch.data(), (int)batch.size(), out);
      {
        std::lock_guard<std::mutex> lk(mu);
        if (!ok) {
          fprintf(stderr, "slave: %s\n", hx.error().c_str());
          failed = stop = true;
        }
        results.insert(results.end(), out.begin(), out.end());
        in_flight = 0;
        batches++;
        blocks_done += batch.size();
      }
      cv.notify_all();
      if (!ok) return;
    }
  });

  // one connection: deliver `send`, receive pairs.  -2 = transport failure (nothing may be assumed delivered unless
  // *delivered is set), otherwise the master's count (-1 = terminate)
  auto exchange = [&](const std::vector<t_result> &send, std::vector<t_pair> &got, bool *delivered) -> int {
    *delivered = false;
    got.clear();
    int fd = socket(AF_INET, SOCK_STREAM, 0);
    if (fd < 0 || connect(fd, (struct sockaddr *)&addr, sizeof(addr)) < 0) {
      if (fd >= 0) close(fd);
      return -2;
    }
    const unsigned int n = (unsigned int)send.size();
    int count = 0;
    bool ok = write_all(fd, &slave_id, sizeof(slave_id)) && write_all(fd, &n, sizeof(n)) &&
              (n == 0 || write_all(fd, send.data(), sizeof(t_result) * n));
    if (ok) *delivered = true;  // the payload is with the master: never send these records again
    ok = ok && read_all(fd, &count, sizeof(count));
    if (ok && count > 0) {
      got.resize((size_t)count);
      ok = read_all(fd, got.data(), sizeof(t_pair) * (size_t)count);
    }
    close(fd);
    if (!ok || count < -1) return -2;  // a corrupt count is a transport error, not a request to spin
    exchanges++;
    return count;
  };

  unsigned int total_targets = 0;
  bool master_done = false;
  std::vector<t_result> send;
  std::vector<t_pair> got;
  while (!master_done) {
    size_t queued, busy;
    {
      std::lock_guard<std::mutex> lk(mu);
      if (failed) break;
      queued = queue.size();
      busy = in_flight;
      send.clear();
      // fetch when the queue runs low; deliver whenever records are waiting
      if (queued < queue_target || !results.empty()) send.swap(results);
    }
    const bool debug_stop = debug_max && total_targets >= debug_max;
    if (debug_stop && queued == 0 && busy == 0 && send.empty()) break;  // everything received has been aligned and delivered
    if (queued >= queue_target && send.empty()) {  // nothing to say to the master right now
      std::unique_lock<std::mutex> lk(mu);
      cv.wait_for(lk, std::chrono::milliseconds(2));
      continue;
    }
    bool delivered = false;
    const int count = exchange(send, got, &delivered);
    if (!delivered && !send.empty()) {  // keep the records for the next connection
      std::lock_guard<std::mutex> lk(mu);
      results.insert(results.begin(), send.begin(), send.end());
    }
    if (count == -2) {
      sleep(1);
      continue;
    }
    if (count == -1) {
      master_done = true;
    } else if (count > 0) {
      {
        std::lock_guard<std::mutex> lk(mu);
        queue.insert(queue.end(), got.begin(), got.end());
      }
      cv.notify_all();
      total_targets += (unsigned int)count;
    } else if (queued == 0 && busy == 0) {
      usleep(50000);  // the master has nothing for us right now
    } else {
      std::unique_lock<std::mutex> lk(mu);  // work is in progress: come back when some of it is done
      cv.wait_for(lk, std::chrono::milliseconds(5));
    }
  }
  {
    std::lock_guard<std::mutex> lk(mu);
    stop = true;
    if (master_done) queue.clear();  // terminate: the master has every result it waits for (Slave.cc:500-502)
  }
  cv.notify_all();
  worker.join();
  printf("== Processing finished ==\nexchanges %llu, GPU batches %llu, blocks %llu (%.1f per batch)\n", exchanges, batches,
         blocks_done, batches ? (double)blocks_done / (double)batches : 0.);
  return failed ? 1 : 0;
}
