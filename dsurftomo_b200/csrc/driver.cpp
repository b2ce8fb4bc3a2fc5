// dsurftomo_b200 -- Fortran-free main program.
//
// Host-only C++ restatement of the reference's main program around the B200 hot path:
//   * DSurfTomo.in, the '#'-gather data file, MOD and MOD.true readers     (src/main.f90:134-335)
//   * the outer inversion loop                                             (src/main.f90:346-592)
//   * writers with the reference's formats: <input>.log, residualFirst.dat, residualLast.dat,
//     <input>Measure.dat.iterNNN, <input>Measure.dat, Vs_model.real, <input>Syn.dat
//                                                                          (src/main.f90:151-166, 396-411, 534-592)
// Everything numerical runs through the C ABI of libdsurf_b200.so (include/dsurftomo_b200.h) and
// stays resident in HBM between CalSurfG and LSMR: plan -> dispersion -> sweeps ->
// dsurf_lsmr_create_from_plan (host glue on the device) -> dsurf_lsmr_solve ->
// dsurf_plan_update_model.  There is no CPU fallback: without an sm_100 device the first ABI call
// fails and the program stops with the library's message.
//
// REAL*4 arithmetic of the host statements (coordinate conversion, delsph, obst = dist/vel,
// mean/std of the residual, grid coordinates of the output files) is restated in float with the
// reference's operand order; build with -ffp-contract=off.
//
//   dsurftomo_b200 [DSurfTomo.in] [--outdir DIR] [--maxiter K] [--seed S] [--raypath] [--quiet]
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/dsurftomo_b200.h"

namespace {

const float kPi = 3.1415926535898f;  // main.f90:57, delsph.f90:4 (REAL*4 parameter)

// delsph.f90:1-28 (REAL*4; sinf/cosf/atan2f/sqrtf = gfortran's REAL*4 intrinsics)
float delsph(float flat1, float flon1, float flat2, float flon2) {
  const float R = 6371.0f;
  const float dlat = flat2 - flat1, dlon = flon2 - flon1;
  const float lat1 = kPi / 2 - flat1, lat2 = kPi / 2 - flat2;
  const float s1 = std::sin(dlat / 2), s2 = std::sin(dlon / 2);
  const float a = s1 * s1 + s2 * s2 * std::cos(lat1) * std::cos(lat2);
  const float c = 2 * std::atan2(std::sqrt(a), std::sqrt(1 - a));
  return R * c;
}

// dnrm2 of lsmrblas.f90:247-315 (scaled sum of squares, REAL*4)
float snrm2(const float *x, int n) {
  if (n < 1) return 0.0f;
  if (n == 1) return std::fabs(x[0]);
  float scale = 0.0f, ssq = 1.0f;
  for (int i = 0; i < n; i++) {
    if (x[i] != 0.0f) {
      const float absxi = std::fabs(x[i]);
      if (scale < absxi) {
        const float r = scale / absxi;
        ssq = 1.0f + ssq * (r * r);
        scale = absxi;
      } else {
        const float r = absxi / scale;
        ssq = ssq + r * r;
      }
    }
  }
  return scale * std::sqrt(ssq);
}

// list-directed tokens: blanks and commas separate, '/' ends the record
std::vector<std::string> tokens(const std::string &line) {
  std::vector<std::string> out;
  std::string cur;
  for (char ch : line) {
    if (ch == '/') break;
    if (ch == ' ' || ch == '\t' || ch == ',' || ch == '\r') {
      if (!cur.empty()) out.push_back(cur), cur.clear();
    } else {
      cur.push_back(ch);
    }
  }
  if (!cur.empty()) out.push_back(cur);
  return out;
}
std::string cnum(std::string s) {  // Fortran D exponents
  for (char &c : s)
    if (c == 'd' || c == 'D') c = 'e';
  return s;
}
bool is_number(const std::string &s0) {
  const std::string s = cnum(s0);
  char *e = nullptr;
  std::strtod(s.c_str(), &e);
  return !s.empty() && e && *e == '\0';
}
double to_num(const std::string &s) { return std::strtod(cnum(s).c_str(), nullptr); }

[[noreturn]] void stop(const std::string &msg) {  // STOP 'msg'
  std::fprintf(stderr, "STOP %s\n", msg.c_str());
  std::exit(2);
}

// gfortran list-directed REAL*4 item (G16.9E2 semantics preceded by one separator blank)
std::string ld_real(float v) {
  char buf[64];
  if (!std::isfinite(v)) {
    std::snprintf(buf, sizeof buf, " %16s", std::isnan(v) ? "NaN" : (v > 0 ? "Infinity" : "-Infinity"));
    return buf;
  }
  int e = 0;
  if (v != 0.0f) {
    std::snprintf(buf, sizeof buf, "%.8E", (double)v);  // exponent after rounding to 9 digits
    e = std::atoi(std::strchr(buf, 'E') + 1);
  }
  if (e >= -1 && e < 9) {
    std::snprintf(buf, sizeof buf, " %12.*f    ", 8 - e, (double)v);
  } else {
    std::snprintf(buf, sizeof buf, " %16.8E", (double)v);
  }
  return buf;
}

struct Tee {  // write(*,...) and write(66,...) pairs
  FILE *log = nullptr;
  bool quiet = false;
  void out(const char *s) const {
    if (!quiet) std::fputs(s, stdout);
  }
  void lg(const char *s) const {
    if (log) std::fputs(s, log);
  }
};

struct Input {
  std::string inputfile, dir, datafile;
  int nx = 0, ny = 0, nz = 0, nsrc = 0, maxiter = 0, ifsyn = 0;
  float goxd = 0, gozd = 0, dvxd = 0, dvzd = 0, weight0 = 0, damp = 0, minthk = 0, minvel = 0, maxvel = 0, spfra = 0;
  float noiselevel = 0, threshold0 = 0;
  int kmaxT[4] = {0, 0, 0, 0};  // Rc, Rg, Lc, Lg
  std::vector<double> t[4];
  int kmax = 0, dall = 0;
  std::vector<float> scxf, sczf, rcxf, rczf, obst, dist, depz, vsf, vsftrue;
  std::vector<int> periods, wavetype, igrt, nrc1, nsrc1;
};

struct LineReader {
  std::ifstream fh;
  std::string name;
  explicit LineReader(const std::string &p) : fh(p), name(p) {}
  bool ok() const { return fh.good(); }
  std::string line() {
    std::string s;
    if (!std::getline(fh, s)) stop("unexpected end of file in " + name);
    return s;
  }
  // read(10,*) a,b,...: keep consuming records until n items have been read
  std::vector<double> nums(size_t n) {
    std::vector<double> v;
    while (v.size() < n) {
      for (const std::string &tk : tokens(line()))
        if (v.size() < n && is_number(tk)) v.push_back(to_num(tk));
    }
    return v;
  }
};

std::string join(const std::string &dir, const std::string &f) {
  if (f.empty() || f[0] == '/' || dir.empty()) return f;
  return dir + "/" + f;
}

// main.f90:134-218: DSurfTomo.in
void read_input(Input &in, const Tee &io) {
  LineReader r(in.inputfile);
  if (!r.ok()) stop("unable to open the inputfile");
  for (int i = 0; i < 3; i++) r.line();
  {
    std::vector<std::string> tk = tokens(r.line());
    if (tk.empty()) stop("data file name missing in " + in.inputfile);
    in.datafile = tk[0];
    if (in.datafile.size() >= 2 && (in.datafile.front() == '\'' || in.datafile.front() == '"'))
      in.datafile = in.datafile.substr(1, in.datafile.size() - 2);
  }
  auto v = r.nums(3);
  in.nx = (int)v[0], in.ny = (int)v[1], in.nz = (int)v[2];
  v = r.nums(2);
  in.goxd = (float)v[0], in.gozd = (float)v[1];
  v = r.nums(2);
  in.dvxd = (float)v[0], in.dvzd = (float)v[1];
  in.nsrc = (int)r.nums(1)[0];
  v = r.nums(2);
  in.weight0 = (float)v[0], in.damp = (float)v[1];
  in.minthk = (float)r.nums(1)[0];
  v = r.nums(2);
  in.minvel = (float)v[0], in.maxvel = (float)v[1];
  in.maxiter = (int)r.nums(1)[0];
  in.spfra = (float)r.nums(1)[0];
  in.kmaxT[0] = (int)r.nums(1)[0];
  char b[256];
  auto both = [&](const char *s) {
    io.out(s);
    io.lg(s);
  };
  io.lg("\n                          S U R F  T O M O\n PLEASE contact Hongjain Fang (fanghj@mail.ustc.edu.cn) if you find any bug\n\n");
  both(" model origin:latitude,longitue\n");
  std::snprintf(b, sizeof b, "%10.5f%10.5f\n", (double)in.goxd, (double)in.gozd);
  both(b);
  both(" grid spacing:latitude,longitue\n");
  std::snprintf(b, sizeof b, "%10.5f%10.5f\n", (double)in.dvxd, (double)in.dvzd);
  both(b);
  both(" model dimension:nx,ny,nz\n");
  std::snprintf(b, sizeof b, "%5d%5d%5d\n", in.nx, in.ny, in.nz);
  both(b);
  static const char *names[4] = {" Rayleigh wave phase velocity used,periods:(s)\n", " Rayleigh wave group velocity used,periods:(s)\n",
                                 " Love wave phase velocity used,periods:(s)\n", " Love wave group velocity used,periods:(s)\n"};
  for (int t = 0; t < 4; t++) {
    if (t > 0) in.kmaxT[t] = (int)r.nums(1)[0];
    if (in.kmaxT[t] > 0) {
      in.t[t] = r.nums(in.kmaxT[t]);
      both(names[t]);
      std::string s;
      for (int i = 0; i < in.kmaxT[t]; i++) {  // (50f7.2)
        std::snprintf(b, sizeof b, "%7.2f", in.t[t][i]);
        s += b;
        if ((i + 1) % 50 == 0 && i + 1 < in.kmaxT[t]) s += "\n";
      }
      s += "\n";
      both(s.c_str());
    }
  }
  in.ifsyn = (int)r.nums(1)[0];
  in.noiselevel = (float)r.nums(1)[0];
  in.threshold0 = (float)r.nums(1)[0];
  in.kmax = in.kmaxT[0] + in.kmaxT[1] + in.kmaxT[2] + in.kmaxT[3];
}

// main.f90:220-283: measurements.  Fortran shapes scxf(nsrc,kmax), rcxf(nrc,nsrc,kmax), ... kept
// column-major, 0-based here.
void read_data(Input &in) {
  const int nsrc = in.nsrc, nrc = in.nsrc, kmax = in.kmax;
  in.scxf.assign((size_t)nsrc * kmax, 0.f);
  in.sczf.assign((size_t)nsrc * kmax, 0.f);
  in.rcxf.assign((size_t)nrc * nsrc * kmax, 0.f);
  in.rczf.assign((size_t)nrc * nsrc * kmax, 0.f);
  in.periods.assign((size_t)nsrc * kmax, 0);
  in.wavetype.assign((size_t)nsrc * kmax, 0);
  in.igrt.assign((size_t)nsrc * kmax, 0);
  in.nrc1.assign((size_t)nsrc * kmax, 0);
  in.nsrc1.assign(kmax, 0);
  std::ifstream fh(join(in.dir, in.datafile));
  if (!fh.good()) stop("unable to open the data file " + in.datafile);
  int istep = 0, istep1 = 0, knumo = 12345, knum = 0;
  float s_lat = 0, s_lon = 0;
  const int kRc = in.kmaxT[0], kRg = in.kmaxT[1], kLc = in.kmaxT[2];
  std::string line;
  while (std::getline(fh, line)) {
    if (line.find_first_not_of(" \t\r") == std::string::npos) continue;
    if (line[0] == '#') {
      std::vector<std::string> tk = tokens(line.substr(1));
      if (tk.size() < 5) stop("bad gather header: " + line);
      float lat = (float)to_num(tk[0]), lon = (float)to_num(tk[1]);
      const int period = (int)to_num(tk[2]), wavetp = (int)to_num(tk[3]), veltp = (int)to_num(tk[4]);
      if (wavetp == 2 && veltp == 0) knum = period;
      if (wavetp == 2 && veltp == 1) knum = kRc + period;
      if (wavetp == 1 && veltp == 0) knum = kRg + kRc + period;
      if (wavetp == 1 && veltp == 1) knum = kLc + kRg + kRc + period;
      if (knum < 1 || knum > kmax) stop("gather header names a period outside DSurfTomo.in: " + line);
      if (knum != knumo) istep = 0;
      istep++;
      istep1 = 0;
      if (istep > nsrc) stop("more gathers per period than nsrc in DSurfTomo.in");
      s_lat = (90.0f - lat) * kPi / 180.0f;
      s_lon = lon * kPi / 180.0f;
      const size_t o = (size_t)(knum - 1) * nsrc + (istep - 1);
      in.scxf[o] = s_lat;
      in.sczf[o] = s_lon;
      in.periods[o] = period;
      in.wavetype[o] = wavetp;
      in.igrt[o] = veltp;
      in.nsrc1[knum - 1] = istep;
      knumo = knum;
    } else {
      std::vector<std::string> tk = tokens(line);
      if (tk.size() < 3) stop("bad measurement line: " + line);
      if (knum < 1) stop("measurement before the first gather header");
      float lat = (float)to_num(tk[0]), lon = (float)to_num(tk[1]);
      const float vel = (float)to_num(tk[2]);
      istep1++;
      if (istep1 > nrc) stop("more receivers per gather than nsrc in DSurfTomo.in");
      lat = (90.0f - lat) * kPi / 180.0f;
      lon = lon * kPi / 180.0f;
      const size_t o = ((size_t)(knum - 1) * nsrc + (istep - 1)) * nrc + (istep1 - 1);
      in.rcxf[o] = lat;
      in.rczf[o] = lon;
      const float d1 = delsph(s_lat, s_lon, lat, lon);
      in.dist.push_back(d1);
      in.obst.push_back(d1 / vel);
      in.nrc1[(size_t)(knum - 1) * nsrc + (istep - 1)] = istep1;
    }
  }
  in.dall = (int)in.obst.size();
}

// main.f90:313-320 (MOD) and :330-336 (MOD.true: no depth line)
void read_model(const std::string &path, int nx, int ny, int nz, std::vector<float> *depz, std::vector<float> &vs) {
  std::ifstream fh(path);
  if (!fh.good()) stop("unable to open " + path);
  std::vector<double> vals;
  std::string line;
  const size_t need = (size_t)nx * ny * nz + (depz ? nz : 0);
  while (vals.size() < need && std::getline(fh, line))
    for (const std::string &tk : tokens(line))
      if (is_number(tk)) vals.push_back(to_num(tk));
  if (vals.size() < need) stop("too few values in " + path);
  size_t o = 0;
  if (depz) {
    depz->resize(nz);
    for (int k = 0; k < nz; k++) (*depz)[k] = (float)vals[o++];
  }
  vs.resize((size_t)nx * ny * nz);
  for (size_t i = 0; i < vs.size(); i++) vs[i] = (float)vals[o++];
}

void check(int rc, const char *what) {
  if (rc != DSURF_OK) {
    std::fprintf(stderr, "%s failed (%d): %s\n", what, rc, dsurf_last_error());
    std::exit(3);
  }
}

// '(5f10.5)' model files: lon, lat, depth, Vs of the interior nodes (main.f90:534-544, 556-590)
void write_model(const std::string &path, const Input &in, const std::vector<float> &vs) {
  FILE *f = std::fopen(path.c_str(), "w");
  if (!f) stop("unable to write " + path);
  for (int k = 1; k <= in.nz - 1; k++)
    for (int j = 1; j <= in.ny - 2; j++)
      for (int i = 1; i <= in.nx - 2; i++) {
        const float lon = in.gozd + (float)(j - 1) * in.dvzd, lat = in.goxd - (float)(i - 1) * in.dvxd;
        std::fprintf(f, "%10.5f%10.5f%10.5f%10.5f\n", (double)lon, (double)lat, (double)in.depz[k - 1],
                     (double)vs[((size_t)(k - 1) * in.ny + j) * in.nx + i]);
      }
  std::fclose(f);
}

void write_residual(const std::string &path, const Input &in, const std::vector<float> &dsyn,
                    const std::vector<float> &dw) {  // main.f90:396-411, list-directed
  FILE *f = std::fopen(path.c_str(), "w");
  if (!f) stop("unable to write " + path);
  for (int i = 0; i < in.dall; i++) {
    std::string s = ld_real(in.dist[i]) + ld_real(dsyn[i]) + ld_real(in.obst[i]) + ld_real(dsyn[i] * dw[i]) +
                    ld_real(in.obst[i] * dw[i]) + ld_real(dw[i]);
    std::fprintf(f, "%s\n", s.c_str());
  }
  std::fclose(f);
}

}  // namespace

int main(int argc, char **argv) {
  Input in;
  Tee io;
  std::string outdir, dump;
  int maxiter_override = -1;
  bool raypath = false;
  uint64_t seed = 20150131ull;
  in.inputfile = "DSurfTomo.in";
  for (int a = 1; a < argc; a++) {
    const std::string s = argv[a];
    if (s == "--outdir" && a + 1 < argc) outdir = argv[++a];
    else if (s == "--maxiter" && a + 1 < argc) maxiter_override = std::atoi(argv[++a]);
    else if (s == "--seed" && a + 1 < argc) seed = std::strtoull(argv[++a], nullptr, 10);
    else if (s == "--quiet") io.quiet = true;
    else if (s == "--raypath") raypath = true;  // raypath.out of the last iteration (CalSurfG.f90:2276-2283)
    else if (s == "--parse-only" && a + 1 < argc) dump = argv[++a];
    else if (s == "--ld-real") {  // formatter self-test: print the list-directed form of each value
      for (int k = a + 1; k < argc; k++) std::printf("[%s]\n", ld_real((float)std::atof(argv[k])).c_str());
      return 0;
    } else in.inputfile = s;
  }
  {
    const size_t p = in.inputfile.find_last_of('/');
    in.dir = p == std::string::npos ? "" : in.inputfile.substr(0, p);
  }
  const std::string base = in.inputfile.substr(in.inputfile.find_last_of('/') == std::string::npos ? 0 : in.inputfile.find_last_of('/') + 1);
  if (outdir.empty()) outdir = in.dir.empty() ? "." : in.dir;
  auto outpath = [&](const std::string &f) { return outdir + "/" + f; };

  io.out("\n                              DSurfTomo (v1.4)\n For bug report, PLEASE contact Hongjain Fang (fanghj1990@gmail.com)\n\n");
  {
    std::ifstream probe(in.inputfile);
    if (!probe.good()) stop("unable to open the inputfile");
  }
  io.log = std::fopen(outpath(base + ".log").c_str(), "w");
  if (!io.log) stop("unable to write the log file in " + outdir);
  read_input(in, io);
  if (maxiter_override >= 0) in.maxiter = maxiter_override;
  read_data(in);
  const int nx = in.nx, ny = in.ny, nz = in.nz, dall = in.dall, kmax = in.kmax, nsrc = in.nsrc;
  // main.f90:287-289 (REAL*4 product truncated to INTEGER)
  const float maxnar_f = in.spfra * (float)dall * (float)nx * (float)ny * (float)nz;
  const long long maxnar = (long long)maxnar_f;
  if (maxnar < 0) io.out(" number overflow, decrease your sparsefrac\n");
  const int maxvp = (nx - 2) * (ny - 2) * (nz - 1);
  char b[512];
  std::snprintf(b, sizeof b, " Number of all measurements%7d\n", dall);
  io.out(b);
  read_model(join(in.dir, "MOD"), nx, ny, nz, &in.depz, in.vsf);
  io.out(" grid points in depth direction:(km)\n");
  {
    std::string s;
    for (int k = 0; k < nz; k++) {
      std::snprintf(b, sizeof b, "%7.2f", (double)in.depz[k]);
      s += b;
    }
    s += "\n";
    io.out(s.c_str());
  }
  const double *tp[4];
  for (int t = 0; t < 4; t++) tp[t] = in.kmaxT[t] > 0 ? in.t[t].data() : nullptr;
  if (!dump.empty()) {  // test hook: parsed inputs as raw little-endian arrays, no device call
    FILE *f = std::fopen(dump.c_str(), "wb");
    if (!f) stop("unable to write " + dump);
    const int hdr[12] = {nx, ny, nz, nsrc, kmax, dall, in.kmaxT[0], in.kmaxT[1], in.kmaxT[2], in.kmaxT[3], in.maxiter, in.ifsyn};
    const float fh[12] = {in.goxd, in.gozd, in.dvxd, in.dvzd, in.weight0, in.damp, in.minthk, in.minvel, in.maxvel, in.spfra,
                          in.noiselevel, in.threshold0};
    std::fwrite(hdr, sizeof hdr, 1, f);
    std::fwrite(fh, sizeof fh, 1, f);
    std::fwrite(&maxnar, sizeof maxnar, 1, f);
    for (int t = 0; t < 4; t++) std::fwrite(in.t[t].data(), sizeof(double), in.t[t].size(), f);
    auto wf = [&](const std::vector<float> &v) { std::fwrite(v.data(), sizeof(float), v.size(), f); };
    auto wi = [&](const std::vector<int> &v) { std::fwrite(v.data(), sizeof(int), v.size(), f); };
    wf(in.scxf), wf(in.sczf), wf(in.rcxf), wf(in.rczf);
    wi(in.periods), wi(in.wavetype), wi(in.igrt), wi(in.nrc1), wi(in.nsrc1);
    wf(in.obst), wf(in.dist), wf(in.depz), wf(in.vsf);
    std::fclose(f);
    std::fclose(io.log);
    return 0;
  }

  // CHECKERBOARD TEST (main.f90:323-343)
  if (in.ifsyn == 1) {
    io.out(" Synthetic Test Begin\n");
    read_model(join(in.dir, "MOD.true"), nx, ny, nz, nullptr, in.vsftrue);
    int rb = 0;
    check(dsurf_synthetic(nx, ny, nz, maxvp, in.vsftrue.data(), in.obst.data(), in.goxd, in.gozd, in.dvxd, in.dvzd,
                          in.kmaxT[0], in.kmaxT[1], in.kmaxT[2], in.kmaxT[3], tp[0], tp[1], tp[2], tp[3],
                          in.wavetype.data(), in.igrt.data(), in.periods.data(), in.depz.data(), in.minthk,
                          in.scxf.data(), in.sczf.data(), in.rcxf.data(), in.rczf.data(), in.nrc1.data(),
                          in.nsrc1.data(), kmax, nsrc, nsrc, in.noiselevel, outdir.c_str(), seed, &rb),
          "synthetic");
  }

  dsurf_plan *plan = nullptr;
  check(dsurf_plan_create(&plan, nx, ny, nz, in.vsf.data(), in.goxd, in.gozd, in.dvxd, in.dvzd, in.kmaxT[0],
                          in.kmaxT[1], in.kmaxT[2], in.kmaxT[3], tp[0], tp[1], tp[2], tp[3], in.wavetype.data(),
                          in.igrt.data(), in.periods.data(), in.depz.data(), in.minthk, in.scxf.data(),
                          in.sczf.data(), in.rcxf.data(), in.rczf.data(), in.nrc1.data(), in.nsrc1.data(), kmax,
                          nsrc, nsrc),
        "plan_create");
  const int ngather = dsurf_plan_num_gathers(plan);
  std::vector<float> dsyn(dall), cbst(dall), dw(dall), dv(maxvp), vs(in.vsf);
  for (int iter = 1; iter <= in.maxiter; iter++) {
    // COMPUTE SENSITIVITY MATRIX (main.f90:353-359): stays in HBM
    io.out(" computing sensitivity matrix...\n");
    check(dsurf_plan_dispersion(plan), "dispersion");
    check(dsurf_plan_reset_rows(plan), "reset_rows");
    if (raypath && iter == in.maxiter) check(dsurf_plan_set_raypath(plan, outpath("raypath.out").c_str(), 0), "set_raypath");
    check(dsurf_plan_sweeps(plan, 0, ngather), "sweeps");
    if (raypath && iter == in.maxiter) check(dsurf_plan_set_raypath(plan, nullptr, 0), "set_raypath");
    int rb = 0;
    check(dsurf_plan_download(plan, nullptr, nullptr, nullptr, dsyn.data(), &rb), "download");
    if (rb) io.out(" Warning: ray paths reached the model boundary (CalSurfG.f90:1447-1454)\n");
    // residual, outlier weights, row scaling, DWS, smoothing rows (main.f90:361-466) on the device
    dsurf_lsmr_sys *sys = nullptr;
    check(dsurf_lsmr_create_from_plan(&sys, plan, in.obst.data(), in.threshold0, in.weight0), "lsmr_create_from_plan");
    float st4[4];
    int m = 0;
    int64_t nar = 0;
    check(dsurf_plan_glue_results(plan, cbst.data(), dw.data(), st4, &m, &nar), "glue_results");
    std::fprintf(io.log, " Maximum and Average DWS values:%s%s\n", ld_real(st4[2]).c_str(), ld_real(st4[3]).c_str());
    if (iter == 1) write_residual(outpath("residualFirst.dat"), in, dsyn, dw);
    if (iter == in.maxiter) write_residual(outpath("residualLast.dat"), in, dsyn, dw);
    if (nar > maxnar) stop("increase sparsity fraction(spfra)");
    // LSMR (main.f90:468-489)
    int istop = 0, itn = 0;
    float anorm = 0, acond = 0, rnorm = 0, arnorm = 0, xnorm = 0;
    check(dsurf_lsmr_solve(sys, in.damp, 1e-6f, 1e-6f, 100.0f, 400, 10, 0, dv.data(), &istop, &itn, &anorm, &acond,
                           &rnorm, &arnorm, &xnorm, nullptr, nullptr, nullptr),
          "lsmr_solve");
    // statistics of the weighted residual (main.f90:491-512), REAL*4 sequential sums
    float sum = 0.0f, sum2 = 0.0f;
    for (int i = 0; i < dall; i++) sum = sum + cbst[i];
    for (int i = 0; i < dall; i++) sum2 = sum2 + cbst[i] * cbst[i];
    const float mean = sum / (float)dall;
    const float std_devs = std::sqrt(sum2 / (float)dall - mean * mean);
    const float rms = snrm2(cbst.data(), dall) / std::sqrt((float)dall);
    std::snprintf(b, sizeof b, "%2dth iteration...\n", iter);
    io.out(b);
    io.lg(b);
    std::snprintf(b, sizeof b, " mean,std_devs and rms of residual after weighting: %8.1fms %8.2fms %8.3f\n",
                  (double)(mean * 1000), (double)(1000 * std_devs), (double)rms);
    io.out(b);
    std::snprintf(b, sizeof b, "mean,std_devs and rms of residual: %8.1fms %8.2fms %8.3f\n", (double)(mean * 1000),
                  (double)(1000 * std_devs), (double)rms);
    io.lg(b);
    const float dvmin = *std::min_element(dv.begin(), dv.end()), dvmax = *std::max_element(dv.begin(), dv.end());
    std::snprintf(b, sizeof b, "min and max velocity variation %7.4f%7.4f\n", (double)dvmin, (double)dvmax);
    io.out(" ");
    io.out(b);
    io.lg(b);
    // model update + clip (main.f90:518-532) on the device; the plan keeps the new model
    check(dsurf_plan_update_model(plan, sys, in.minvel, in.maxvel, nullptr, vs.data()), "update_model");
    check(dsurf_lsmr_destroy(sys), "lsmr_destroy");
    std::snprintf(b, sizeof b, "%sMeasure.dat.iter%03d", base.c_str(), iter);
    write_model(outpath(b), in, vs);
  }
  io.out(" Program finishes successfully\n");
  io.lg(" Program finishes successfully\n");
  if (in.ifsyn == 1) {
    write_model(outpath("Vs_model.real"), in, in.vsftrue);
    write_model(outpath(base + "Syn.dat"), in, vs);
    const std::string msg = " Output True velocity model to Vs_model.real\n Output inverted shear velocity model to " + base + "Syn.dat\n";
    io.out(msg.c_str());
    io.lg(msg.c_str());
  } else {
    write_model(outpath(base + "Measure.dat"), in, vs);
    const std::string msg = " Output inverted shear velocity model to " + base + "Measure.dat\n";
    io.out(msg.c_str());
    io.lg(msg.c_str());
  }
  std::fclose(io.log);
  dsurf_plan_destroy(plan);
  return 0;
}
