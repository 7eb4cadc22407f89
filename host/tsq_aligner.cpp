// tsq_aligner.cpp -- the B200 backend behind the boundary tweakseq ALREADY has: a child process.
//
// tweakseq reaches its aligner through QProcess::start(exec, args) (UI/SeqEditMainWin.cpp:1654-1660) with
// the argv its tool wrapper builds:
//   ClustalO  Core/ClustalO.cpp:51   --force -v --outfmt=fa --output-order=tree-order -i <fin> -o <fout>
//   MUSCLE    Core/Muscle.cpp:52     -in <fin> -out <fout>
//   MAFFT     Core/MAFFT.cpp:52,98   --auto --thread -1 <fin>          (alignment on stdout)
// and asks for the version with `--version` (ClustalO.cpp:100-111, MAFFT.cpp:107) or `-version`
// (Muscle.cpp:103).  This program accepts all three conventions and runs libtsqb200.so (tsq_run_fasta with
// TSQ_FLAG_MSA_OUT), so an UNMODIFIED tweakseq uses the B200 backend by pointing the tool's path setting
// (<alignment_tool><path>, ClustalO.cpp:63-86) at this binary; the in-process adapter of host/qt/ is the
// other way in.  stdout/stderr lines reach MessageWin as for any tool (SeqEditMainWin.cpp:822-834); exit
// code 0 = success (:836-861).  "Stop" is QProcess::kill(): nothing to clean up, the driver reclaims the GPU.
// No CPU fallback: without a B200 it fails with a message and exit code 1.
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include <unistd.h>

#include "../include/tsq_b200.h"

namespace {

struct Options {
  std::string in, out;
  bool to_stdout = false, version = false, dry_run = false, verbose = false, keep_distmat = false, matrix_only = false;
  int alphabet = -1;   // -1 = detect from the residues
  int gap_open = -1, gap_extend = -1, device = 0;
  int devices = 1;   // --devices N | all: B200s of the box this job is spread over (tsq_params.n_devices)
  bool identity = false;
  bool kimura = false;        // --kimura: Kimura-corrected identity distances (implies --identity-distance)
  bool input_order = false;   // clustalo --output-order=input-order (default: tree order, what tweakseq asks for)
};

void usage(FILE* f) {
  fputs("usage: tsq-aligner -i IN.fa -o OUT.fa            (clustalo style; --in/--out, --infile/--outfile too)\n"
        "       tsq-aligner -in IN.fa -out OUT.fa        (muscle style)\n"
        "       tsq-aligner [--auto --thread N] IN.fa    (mafft style: alignment on stdout)\n"
        "       tsq-aligner --version | -version\n"
        "options: --seqtype=Protein|DNA|RNA, --amino, --nuc (default: detected), --gap-open N, --gap-extend N,\n"
        "         --device N, --devices N|all (several B200s; env TSQ_DEVICES), --identity-distance, --kimura (corrected identity distance; fails above D = 0.75), --distmat-out (keep OUT.fa.distmat), --matrix-only (OUT = matrix),\n"
        "         --dry-run (print the parsed job and exit); clustalo's --force -v --outfmt=fa --output-order=... are accepted\n",
        f);
}

// exactly --name or --name=value (prefix matching would take clustalo's --infmt or mafft's --inputorder for --in)
bool is_opt(const std::string& s, const char* name) {
  const size_t l = strlen(name);
  return s.compare(0, l, name) == 0 && (s.size() == l || s[l] == '=');
}

// Returns 0 = ok, 2 = usage error (message in err).
int parse(int argc, char** argv, Options& o, std::string& err) {
  std::vector<std::string> positional;
  for (int k = 1; k < argc; k++) {
    const std::string a = argv[k];
    auto value = [&](std::string& dst) -> bool {   // "--opt value" or "--opt=value"
      const size_t eq = a.find('=');
      if (eq != std::string::npos) { dst = a.substr(eq + 1); return true; }
      if (k + 1 >= argc) { err = "option " + a + " needs a value"; return false; }
      dst = argv[++k];
      return true;
    };
    std::string v;
    if (a == "--version" || a == "-version") o.version = true;
    else if (a == "-h" || a == "--help") { err = "help"; return 2; }
    else if (a == "-i" || a == "-in" || is_opt(a, "--in") || is_opt(a, "--infile")) { if (!value(o.in)) return 2; }
    else if (a == "-o" || a == "-out" || is_opt(a, "--out") || is_opt(a, "--outfile")) { if (!value(o.out)) return 2; }
    else if (is_opt(a, "--infmt")) {
      if (!value(v)) return 2;
      if (v != "fa" && v != "fasta" && v != "a2m") { err = "only FASTA input is read (--infmt=" + v + ")"; return 2; }
    }
    else if (a == "--inputorder" || a == "--reorder") {}                    // mafft: row order; tree order is what tweakseq asks for
    else if (is_opt(a, "--outfmt")) {
      if (!value(v)) return 2;
      if (v != "fa" && v != "fasta" && v != "a2m") { err = "only FASTA output is produced (--outfmt=" + v + ")"; return 2; }
    }
    else if (is_opt(a, "--output-order")) {
      if (!value(v)) return 2;
      if (v == "input-order") o.input_order = true;
      else if (v != "tree-order") { err = "unknown --output-order " + v; return 2; }
    }
    else if (is_opt(a, "--seqtype") || a == "-t") {
      if (!value(v)) return 2;
      for (char& ch : v) ch = (char)tolower((unsigned char)ch);
      if (v == "protein") o.alphabet = TSQ_PROTEIN;
      else if (v == "dna" || v == "rna") o.alphabet = TSQ_NUCLEOTIDE;
      else { err = "unknown --seqtype " + v; return 2; }
    }
    else if (a == "--amino") o.alphabet = TSQ_PROTEIN;
    else if (a == "--nuc") o.alphabet = TSQ_NUCLEOTIDE;
    else if (is_opt(a, "--gap-open")) { if (!value(v)) return 2; o.gap_open = atoi(v.c_str()); }
    else if (is_opt(a, "--gap-extend")) { if (!value(v)) return 2; o.gap_extend = atoi(v.c_str()); }
    else if (is_opt(a, "--device")) { if (!value(v)) return 2; o.device = atoi(v.c_str()); }
    else if (is_opt(a, "--devices")) { if (!value(v)) return 2; o.devices = v == "all" ? -1 : atoi(v.c_str()); }
    else if (is_opt(a, "--thread") || is_opt(a, "--threads")) { if (!value(v)) return 2; }       // mafft: host threads mean nothing here
    else if (a == "--identity-distance") o.identity = true;
    else if (a == "--kimura") o.identity = o.kimura = true;
    else if (a == "--distmat-out") o.keep_distmat = true;
    else if (a == "--matrix-only") o.matrix_only = true;
    else if (a == "--dry-run") o.dry_run = true;
    else if (a == "-v" || a == "--verbose") o.verbose = true;
    else if (a == "--force" || a == "--auto" || a == "--quiet") {}          // nothing to force, tune or silence
    else if (!a.empty() && a[0] == '-' && a != "-") { err = "unknown option " + a; return 2; }
    else positional.push_back(a);
  }
  if (const char* e = getenv("TSQ_DEVICES"))   // an unmodified tweakseq cannot add options to the wrapper's argv
    if (o.devices == 1) o.devices = strcmp(e, "all") == 0 ? -1 : atoi(e);
  if (o.version) return 0;
  if (o.in.empty() && positional.size() == 1) { o.in = positional[0]; positional.clear(); }   // mafft style
  if (!positional.empty()) { err = "unexpected argument " + positional[0]; return 2; }
  if (o.in.empty()) { err = "no input file"; return 2; }
  if (o.out.empty()) o.to_stdout = true;
  return 0;
}

// ClustalO decides protein vs nucleotide from the residues when --seqtype is absent; so does this: at
// least 90 % of the letters outside header lines being ACGTUN means nucleotide.
int detect_alphabet(const std::string& path) {
  std::ifstream in(path);
  std::string line;
  unsigned long long letters = 0, nuc = 0;
  while (std::getline(in, line)) {
    if (!line.empty() && (line[0] == '>' || line[0] == ';')) continue;
    for (unsigned char ch : line) {
      if (!isalpha(ch)) continue;
      letters++;
      if (strchr("ACGTUNacgtun", ch)) nuc++;
    }
  }
  return (letters > 0 && nuc * 10 >= letters * 9) ? TSQ_NUCLEOTIDE : TSQ_PROTEIN;
}

void log_line(void* user, const char* line) {
  // with the alignment on stdout (mafft style) progress goes to stderr; tweakseq shows both (SeqEditMainWin.cpp:822-834)
  FILE* f = *static_cast<bool*>(user) ? stderr : stdout;
  fprintf(f, "%s\n", line);
  fflush(f);
}

}  // namespace

int main(int argc, char** argv) {
  Options o;
  std::string err;
  const int prc = parse(argc, argv, o, err);
  if (prc != 0) {
    if (err != "help") fprintf(stderr, "tsq-aligner: %s\n", err.c_str());
    usage(err == "help" ? stdout : stderr);
    return err == "help" ? 0 : 2;
  }
  if (o.version) {
    // what AlignmentTool::version() then shows: ClustalO and MUSCLE read the probe's stdout (ClustalO.cpp:107,
    // Muscle.cpp:107, which keeps the second word), MAFFT its stderr (MAFFT.cpp:111) -- so both get the line
    printf("%s\n", tsq_version_string());
    fprintf(stderr, "%s\n", tsq_version_string());
    return 0;
  }
  {
    std::ifstream probe(o.in);
    if (!probe) {
      fprintf(stderr, "tsq-aligner: cannot open %s\n", o.in.c_str());
      return 1;
    }
  }
  const int alphabet = o.alphabet >= 0 ? o.alphabet : detect_alphabet(o.in);
  std::string out = o.out;
  if (o.to_stdout) {
    char tmpl[] = "/tmp/tsq-aligner.XXXXXX";
    const int fd = o.dry_run ? -1 : mkstemp(tmpl);
    if (!o.dry_run && fd < 0) {
      fprintf(stderr, "tsq-aligner: cannot create a temporary file\n");
      return 1;
    }
    if (fd >= 0) close(fd);
    out = tmpl;
  }
  if (o.dry_run) {
    printf("in=%s out=%s alphabet=%s gap_open=%d gap_extend=%d device=%d devices=%d output=%s order=%s\n", o.in.c_str(),
           o.to_stdout ? "<stdout>" : out.c_str(), alphabet == TSQ_NUCLEOTIDE ? "nucleotide" : "protein", o.gap_open, o.gap_extend,
           o.device, o.devices, o.matrix_only ? "matrix" : "alignment", o.input_order ? "input" : "tree");
    return 0;
  }
  tsq_params p;
  tsq_default_params(&p);
  p.alphabet = alphabet;
  p.gap_open = o.gap_open;
  p.gap_extend = o.gap_extend;
  p.device = o.device;
  p.n_devices = o.devices;
  if (!o.matrix_only) p.flags |= TSQ_FLAG_MSA_OUT;
  if (o.keep_distmat) p.flags |= TSQ_FLAG_KEEP_DISTMAT;
  if (o.identity) p.flags |= TSQ_FLAG_IDENTITY;
  if (o.kimura) p.flags |= TSQ_FLAG_KIMURA;
  if (o.input_order) p.flags |= TSQ_FLAG_INPUT_ORDER;
  bool to_stderr = o.to_stdout;
  const int rc = tsq_run_fasta(o.in.c_str(), out.c_str(), &p, log_line, &to_stderr, nullptr);
  if (rc != TSQ_OK) {
    fprintf(stderr, "tsq-aligner: %s%s\n", tsq_status_string(rc),
            rc == TSQ_ERR_NO_DEVICE ? " (this aligner runs on a B200 only; there is no CPU path)" : "");
    if (o.to_stdout) unlink(out.c_str());
    return 1;
  }
  if (o.to_stdout) {
    std::ifstream res(out, std::ios::binary);
    char buf[1 << 16];
    while (res.read(buf, sizeof buf) || res.gcount() > 0) fwrite(buf, 1, (size_t)res.gcount(), stdout);
    fflush(stdout);
    unlink(out.c_str());
    unlink((out + ".dnd").c_str());
    if (o.keep_distmat) unlink((out + ".distmat").c_str());
  }
  return 0;
}
