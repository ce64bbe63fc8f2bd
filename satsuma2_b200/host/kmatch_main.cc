// KMatch on B200: drop-in for the reference's seeding program (kmatch/KMatch.cc:320-343), same positional arguments
//   KMatch query.fa target.fa K output min_length jump max_freq        [-device d as an optional 8th/9th argument]
// and the same output: raw t_result records (prob = ident = 1) that SatsumaSynteny2 loads as seeds
// (analysis/SatsumaSynteny2.cc:415-434).  FASTA reading follows the reference's own loop (KMatch.cc:37-41, 113-123):
// a line starting with '>' opens a record, every other line is appended as it is (no case folding, no trimming).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "../../include/satsuma_kmatch.h"

struct Fasta {
  std::string blob;
  std::vector<int64_t> offsets, lens;
};

// same records as the reference's getline loop (KMatch.cc:37-41, 113-123), read in one piece
static bool read_fasta_raw(const char *path, Fasta &f) {
  FILE *fp = fopen(path, "rb");
  if (!fp) return false;
  fseek(fp, 0, SEEK_END);
  const long size = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  std::string text((size_t)(size > 0 ? size : 0), '\0');
  const size_t got = size > 0 ? fread(&text[0], 1, (size_t)size, fp) : 0;
  fclose(fp);
  text.resize(got);
  f.blob.reserve(got);
  size_t rec_start = 0;  // offset in blob where the open record began
  bool open = false;
  auto close_record = [&]() {
    if (open && f.blob.size() > rec_start) {  // empty records are skipped, as in the reference
      f.offsets.push_back((int64_t)rec_start);
      f.lens.push_back((int64_t)(f.blob.size() - rec_start));
    }
    rec_start = f.blob.size();
  };
  size_t pos = 0;
  while (pos < text.size()) {
    const char *nl = (const char *)memchr(text.data() + pos, '\n', text.size() - pos);
    const size_t end = nl ? (size_t)(nl - text.data()) : text.size();
    if (end > pos && text[pos] == '>') {
      close_record();
      open = true;
    } else if (end > pos) {
      if (!open) {  // sequence lines before any header: the reference collects them into a record as well
        open = true;
        rec_start = f.blob.size();
      }
      f.blob.append(text, pos, end - pos);
    }
    pos = end + 1;
  }
  close_record();
  return true;
}

int main(int argc, char **argv) {
  if (argc != 8 && argc != 10) {
    printf("Usage: %s query.fa target.fa K output.fa min_length jump max_freq [-device d]\n", argv[0]);
    return -1;
  }
  if (atoi(argv[3]) % 2 == 0) {
    printf("KMatch only accepts odd K values, please try again\n");
    return 1;
  }
  sx_kmatch_config cfg;
  sx_kmatch_default_config(&cfg);
  cfg.k = atoi(argv[3]);
  cfg.min_length = atoi(argv[5]);
  cfg.max_jump = atoi(argv[6]);
  cfg.max_freq = atoi(argv[7]);
  if (argc == 10 && strcmp(argv[8], "-device") == 0) cfg.device = atoi(argv[9]);
  Fasta q, t;
  if (!read_fasta_raw(argv[1], q) || !read_fasta_raw(argv[2], t)) {
    fprintf(stderr, "KMatch(B200): cannot read the FASTA files\n");
    return 1;
  }
  std::vector<sx_result> out((size_t)1 << 16);
  int64_t n = 0;
  sx_kmatch_stats st;
  auto run = [&]() {
    return sx_kmatch(&cfg, q.blob.data(), q.offsets.data(), q.lens.data(), (int32_t)q.lens.size(), t.blob.data(),
                     t.offsets.data(), t.lens.data(), (int32_t)t.lens.size(), out.data(), (int64_t)out.size(), &n, &st);
  };
  int rc = run();
  if (rc == SX_ERR_CAPACITY) {
    out.resize((size_t)n);
    rc = run();
  }
  if (rc != SX_OK) {
    fprintf(stderr, "KMatch(B200): %s\n", sx_kmatch_last_error());
    return 1;
  }
  printf("Kmer arrays filtered to %lld (query) / %lld (target) elements\n%lld matching positions\n", (long long)st.query_kmers,
         (long long)st.target_kmers, (long long)st.kmer_matches);
  FILE *fo = fopen(argv[4], "wb");
  if (!fo || (n > 0 && fwrite(out.data(), sizeof(sx_result), (size_t)n, fo) != (size_t)n)) {
    fprintf(stderr, "KMatch(B200): cannot write %s\n", argv[4]);
    return 1;
  }
  fclose(fo);
  printf("%lld matches dumped\ndevice work (upload, kernels, download): %.1f ms\n", (long long)n, st.gpu_ms);
  return 0;
}
