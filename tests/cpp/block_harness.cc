// TEST INFRASTRUCTURE: drives the C++ GPU blocks of include/gr_b200_blocks.h the way the reference's
// QA code drives its blocks (gr_vector_source -> block -> gr_vector_sink under a single-threaded
// scheduler that honours history(), forecast(), output_multiple() and consume()), reading and
// writing raw binary files so that tests/test_cpp_blocks.py can compare against the oracle.
//
//   block_harness run <block spec...> <in.bin> <out.bin> <max_noutput>
//   block_harness errors      argument-error -> exception-type mapping (no GPU needed)
//   block_harness contract    scheduler-visible contracts (needs a GPU)
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "gr_b200_blocks.h"

using namespace gr_b200;

static std::vector<char> slurp(const char* path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
  return std::vector<char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
static std::vector<float> floats(const char* path) {
  if (!strcmp(path, "-")) return std::vector<float>();
  std::vector<char> b = slurp(path);
  return std::vector<float>((const float*)b.data(), (const float*)(b.data() + b.size() / 4 * 4));
}

// vector_source -> blk -> vector_sink.  `streams`: the input streams (each WITHOUT history); the
// harness pre-loads history()-1 zero items like gr_buffer does (gr_buffer.cc:201-214).
static std::vector<char> run_block(gr_block& blk, const std::vector<std::vector<char> >& streams, int max_noutput) {
  const int nin = (int)streams.size();
  const size_t isz = blk.input_signature()->sizeof_stream_item(0), osz = blk.output_signature()->sizeof_stream_item(0);
  const size_t nitems = streams[0].size() / isz;
  std::vector<std::vector<char> > buf(nin);
  size_t hist = blk.history();
  auto rebuild = [&](size_t consumed_total) {  // history may change after set_taps: keep buffers = [hist-1 old items][rest]
    for (int i = 0; i < nin; i++) {
      std::vector<char> nb((hist - 1) * isz + streams[i].size(), 0);
      memcpy(nb.data() + (hist - 1) * isz, streams[i].data(), streams[i].size());
      buf[i].swap(nb);
    }
    (void)consumed_total;
  };
  rebuild(0);
  std::vector<char> out;
  std::vector<char> obuf((size_t)max_noutput * osz);
  size_t rp = 0;  // items consumed so far (read pointer = rp, pointing at the first history item)
  int idle = 0;
  while (true) {
    if (blk.history() != hist) { hist = blk.history(); rebuild(rp); }
    const long avail = (long)nitems - (long)rp;  // new items not yet consumed
    int noutput = max_noutput - max_noutput % blk.output_multiple();
    gr_vector_int req(nin);
    for (; noutput > 0; noutput -= blk.output_multiple()) {
      blk.forecast(noutput, req);
      if ((long)req[0] <= avail + (long)hist - 1) break;
    }
    if (noutput <= 0) break;
    gr_vector_int ninput(nin, (int)(avail + hist - 1));
    gr_vector_const_void_star in(nin);
    for (int i = 0; i < nin; i++) in[i] = buf[i].data() + rp * isz;
    gr_vector_void_star outv(1, obuf.data());
    blk.b200_reset_consumed();
    const int r = blk.general_work(noutput, ninput, in, outv);
    if (r == gr_block::WORK_DONE) break;
    const int c = blk.b200_consumed(0);
    if (r > 0) out.insert(out.end(), obuf.data(), obuf.data() + (size_t)r * osz);
    rp += c;
    if (r == 0 && c == 0) { if (++idle > 2) break; } else idle = 0;
  }
  return out;
}

static int cmd_run(int argc, char** argv) {
  // argv: run <kind> <args...> in out max_noutput
  const std::string kind = argv[2];
  const char* inpath = argv[argc - 3];
  const char* outpath = argv[argc - 2];
  const int max_noutput = atoi(argv[argc - 1]);
  std::vector<char> in = slurp(inpath);
  std::vector<std::vector<char> > streams(1, in);
  std::vector<char> out;
  if (kind == "fir_ccf") {
    auto b = gr_make_fir_filter_ccf(atoi(argv[3]), floats(argv[4]));
    out = run_block(*b, streams, max_noutput);
  } else if (kind == "fir_fff") {
    auto b = gr_make_fir_filter_fff(atoi(argv[3]), floats(argv[4]));
    out = run_block(*b, streams, max_noutput);
  } else if (kind == "fxlat") {
    auto b = gr_make_freq_xlating_fir_filter_ccf(atoi(argv[3]), floats(argv[4]), atof(argv[5]), atof(argv[6]));
    out = run_block(*b, streams, max_noutput);
  } else if (kind == "pfb") {
    const unsigned M = (unsigned)atoi(argv[3]);
    auto b = gr_make_pfb_channelizer_ccf(M, floats(argv[4]), (float)atof(argv[5]));
    // gr_stream_to_streams: stream j = x[m*M + j]  (general/gr_stream_to_streams.cc:57-63)
    const size_t n = in.size() / sizeof(gr_complex) / M;
    streams.assign(M, std::vector<char>(n * sizeof(gr_complex)));
    for (size_t m = 0; m < n; m++)
      for (unsigned j = 0; j < M; j++)
        memcpy(streams[j].data() + m * sizeof(gr_complex), in.data() + (m * M + j) * sizeof(gr_complex), sizeof(gr_complex));
    out = run_block(*b, streams, max_noutput);
  } else if (kind == "pfbdec") {
    const unsigned M = (unsigned)atoi(argv[3]);
    auto b = gr_make_pfb_decimator_ccf(M, floats(argv[4]), (unsigned)atoi(argv[5]));
    const size_t n = in.size() / sizeof(gr_complex) / M;
    streams.assign(M, std::vector<char>(n * sizeof(gr_complex)));
    for (size_t m = 0; m < n; m++)
      for (unsigned j = 0; j < M; j++)
        memcpy(streams[j].data() + m * sizeof(gr_complex), in.data() + (m * M + j) * sizeof(gr_complex), sizeof(gr_complex));
    out = run_block(*b, streams, max_noutput);
  } else if (kind == "fftfilt") {
    std::vector<float> t = floats(argv[4]);  // interleaved re, im
    std::vector<gr_complex> tc(t.size() / 2);
    for (size_t i = 0; i < tc.size(); i++) tc[i] = gr_complex(t[2 * i], t[2 * i + 1]);
    auto b = gr_make_fft_filter_ccc(atoi(argv[3]), tc);
    out = run_block(*b, streams, max_noutput);
  } else if (kind == "arb") {
    auto b = gr_make_pfb_arb_resampler_ccf((float)atof(argv[3]), floats(argv[4]), (unsigned)atoi(argv[5]));
    out = run_block(*b, streams, max_noutput);
  } else if (kind == "fft") {
    auto b = gr_make_fft_vcc(atoi(argv[3]), atoi(argv[4]) != 0, floats(argv[5]), atoi(argv[6]) != 0);
    out = run_block(*b, streams, max_noutput);
  } else if (kind == "quad") {
    auto b = gr_make_quadrature_demod_cf((float)atof(argv[3]));
    out = run_block(*b, streams, max_noutput);
  } else if (kind == "mm") {
    auto b = digital_make_clock_recovery_mm_ff((float)atof(argv[3]), (float)atof(argv[4]), (float)atof(argv[5]), (float)atof(argv[6]),
                                               (float)atof(argv[7]));
    out = run_block(*b, streams, max_noutput);
    fprintf(stderr, "mm final mu %.9g omega %.9g\n", b->mu(), b->omega());
  } else if (kind == "slicer4") {
    auto b = pager_make_slicer_fb((float)atof(argv[3]));
    out = run_block(*b, streams, max_noutput);
  } else if (kind == "slicer2") {
    auto b = digital_make_binary_slicer_fb();
    out = run_block(*b, streams, max_noutput);
  } else if (kind == "corr") {
    auto b = digital_make_correlate_access_code_bb(argv[3], atoi(argv[4]));
    out = run_block(*b, streams, max_noutput);
  } else if (kind == "map") {
    std::vector<float> m = floats(argv[3]);
    auto b = gr_make_map_bb(std::vector<int>(m.begin(), m.end()));
    out = run_block(*b, streams, max_noutput);
  } else if (kind == "unpack") {
    auto b = gr_make_unpack_k_bits_bb((unsigned)atoi(argv[3]));
    out = run_block(*b, streams, max_noutput);
  } else if (kind == "mmcc") {
    auto b = digital_make_clock_recovery_mm_cc((float)atof(argv[3]), (float)atof(argv[4]), (float)atof(argv[5]), (float)atof(argv[6]),
                                               (float)atof(argv[7]));
    out = run_block(*b, streams, max_noutput);
    fprintf(stderr, "mmcc final mu %.9g omega %.9g\n", b->mu(), b->omega());
  } else if (kind == "framer") {
    // a sink: the scheduler hands it whatever is there, max_noutput items at a time; the messages are what comes out
    gr_msg_queue_sptr q = gr_make_msg_queue();
    auto b = gr_make_framer_sink_1(q);
    gr_vector_void_star none;
    for (size_t pos = 0; pos < in.size(); pos += (size_t)max_noutput) {
      const int n = (int)std::min<size_t>((size_t)max_noutput, in.size() - pos);
      gr_vector_const_void_star iv(1, in.data() + pos);
      if (b->work(n, iv, none) != n) return 3;
    }
    while (q->count()) {   // record: arg1 (1 byte), length (2 bytes, little endian), payload
      gr_message_sptr m = q->delete_head_nowait();
      const unsigned len = (unsigned)m->length();
      out.push_back((char)(int)m->arg1());
      out.push_back((char)(len & 255));
      out.push_back((char)(len >> 8));
      out.insert(out.end(), (const char*)m->msg(), (const char*)m->msg() + len);
    }
  } else if (kind == "s2s" || kind == "v2s") {
    const size_t isz = (size_t)atoi(argv[3]), ns = (size_t)atoi(argv[4]);
    const size_t n = in.size() / isz / ns;
    std::vector<std::vector<char> > outs(ns, std::vector<char>(n * isz));
    gr_vector_const_void_star iv(1);
    gr_vector_void_star ov(ns);
    for (size_t pos = 0; pos < n; pos += (size_t)max_noutput) {
      const int k = (int)std::min<size_t>((size_t)max_noutput, n - pos);
      iv[0] = in.data() + pos * ns * isz;
      for (size_t j = 0; j < ns; j++) ov[j] = outs[j].data() + pos * isz;
      int r;
      if (kind == "s2s") { auto b = gr_make_stream_to_streams(isz, ns); r = b->work(k, iv, ov); }
      else { auto b = gr_make_vector_to_streams(isz, ns); r = b->work(k, iv, ov); }
      if (r != k) return 3;
    }
    for (size_t j = 0; j < ns; j++) out.insert(out.end(), outs[j].begin(), outs[j].end());   // stream after stream
  } else {
    fprintf(stderr, "unknown block %s\n", kind.c_str());
    return 2;
  }
  std::ofstream f(outpath, std::ios::binary);
  f.write(out.data(), (std::streamsize)out.size());
  return 0;
}

template <class F> static const char* what_throws(F f) {
  try { f(); } catch (const std::invalid_argument&) { return "invalid_argument"; } catch (const std::out_of_range&) { return "out_of_range"; }
  catch (const std::runtime_error&) { return "runtime_error"; } catch (...) { return "other"; }
  return "none";
}

static int cmd_errors() {
  std::vector<float> t(8, 1.f);
  // the reference throws these from the constructors (SURVEY.md 8b "Errors"); argument checks come
  // before the device is touched, so the mapping is testable without a GPU
  printf("pfb_bad_oversample %s\n", what_throws([&] { gr_make_pfb_channelizer_ccf(8, t, 3.0f); }));       // invalid_argument (:57-60)
  printf("mm_omega_lt_1 %s\n", what_throws([&] { digital_make_clock_recovery_mm_ff(0.5f, 0.1f, 0.5f, 0.1f, 0.001f); }));  // out_of_range
  printf("mm_negative_gain %s\n", what_throws([&] { digital_make_clock_recovery_mm_ff(2.f, -0.1f, 0.5f, 0.1f, 0.001f); }));
  printf("corr_code_too_long %s\n", what_throws([&] { digital_make_correlate_access_code_bb(std::string(65, '1'), 0); }));  // out_of_range
  printf("fft_size_zero %s\n", what_throws([&] { gr_make_fft_vcc(0, true, std::vector<float>(), false); }));   // out_of_range (gri_fft.cc:104-105)
  printf("io_signature %s\n", what_throws([&] { gr_make_io_signature(2, 1, 4); }));
  printf("unpack_k_zero %s\n", what_throws([&] { gr_make_unpack_k_bits_bb(0); }));                           // out_of_range (:44-45)
  printf("mmcc_omega_zero %s\n", what_throws([&] { digital_make_clock_recovery_mm_cc(0.f, 0.1f, 0.5f, 0.1f, 0.001f); }));  // out_of_range (:65-66)
  printf("mmcc_negative_gain %s\n", what_throws([&] { digital_make_clock_recovery_mm_cc(2.f, 0.1f, 0.5f, -0.1f, 0.001f); }));
  return 0;
}

static int cmd_contract() {
  int bad = 0;
#define EXPECT(cond) do { if (!(cond)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); bad++; } else printf("ok %s\n", #cond); } while (0)
  std::vector<float> t(12, 0.25f);
  {
    auto b = gr_make_fir_filter_ccf(3, t);
    EXPECT(b->history() == 12 && b->decimation() == 3 && b->relative_rate() == 1.0 / 3 && b->fixed_rate());
    gr_vector_int req(1);
    b->forecast(10, req);
    EXPECT(req[0] == 10 * 3 + 11);
    std::vector<gr_complex> in(200, gr_complex(1, -1)), out(50);
    gr_vector_const_void_star iv(1, in.data());
    gr_vector_void_star ov(1, out.data());
    gr_vector_int ni(1, 200);
    EXPECT(b->general_work(20, ni, iv, ov) == 20 && b->b200_consumed(0) == 60);
    EXPECT(std::abs(out[5] - gr_complex(3, -3)) < 1e-5f);
    b->set_taps(std::vector<float>(5, 1.f));
    b->b200_reset_consumed();
    EXPECT(b->general_work(20, ni, iv, ov) == 0 && b->b200_consumed(0) == 0);   // gr_fir_filter_XXX.cc.t:74-79
    EXPECT(b->history() == 5);
    EXPECT(b->general_work(20, ni, iv, ov) == 20 && std::abs(out[5] - gr_complex(5, -5)) < 1e-5f);
  }
  {
    const unsigned M = 8;
    auto b = gr_make_pfb_channelizer_ccf(M, std::vector<float>(32, 1.f / 32), 1);
    EXPECT(b->history() == 5 && b->output_multiple() == 1 && b->relative_rate() == 1.0 / 8);
    auto b2 = gr_make_pfb_channelizer_ccf(M, std::vector<float>(32, 1.f / 32), 2);
    EXPECT(b2->output_multiple() == 2 && b2->relative_rate() == 0.25);
    std::vector<std::vector<gr_complex> > s(M, std::vector<gr_complex>(64, gr_complex(1, 0)));
    gr_vector_const_void_star iv(M);
    for (unsigned j = 0; j < M; j++) iv[j] = s[j].data();
    std::vector<gr_complex> out(M * 16);
    gr_vector_void_star ov(1, out.data());
    gr_vector_int ni(M, 64);
    EXPECT(b->general_work(16, ni, iv, ov) == 0);                             // gr_pfb_channelizer_ccf.cc:164-167
    b->b200_reset_consumed();
    EXPECT(b->general_work(16, ni, iv, ov) == 16 && b->b200_consumed(0) == 16 && b->b200_consumed(M - 1) == 16);
    EXPECT(std::abs(out[0] - gr_complex(1, 0)) < 1e-5f && std::abs(out[1]) < 1e-5f);  // DC lands in bin 0
  }
  {
    auto b = digital_make_clock_recovery_mm_ff(2.f, 0.01f, 0.5f, 0.01f, 0.001f);
    gr_vector_int req(1);
    b->forecast(10, req);
    EXPECT(req[0] == 28);                                                      // qa_clock_recovery_mm.py / :80-87
    EXPECT(b->mu() == 0.5f && b->omega() == 2.f && b->gain_mu() == 0.01f && b->relative_rate() == 0.5);
    b->set_omega(2.5f);
    EXPECT(b->omega() == 2.5f);
  }
  {
    auto b = gr_make_fft_vcc(64, true, std::vector<float>(), false);
    EXPECT(b->set_window(std::vector<float>(64, 1.f)) && !b->set_window(std::vector<float>(63, 1.f)));
    auto q = gr_make_quadrature_demod_cf(2.5f);
    EXPECT(q->history() == 2 && q->gain() == 2.5f);
    auto c = digital_make_correlate_access_code_bb("1011", 0);
    EXPECT(c->set_access_code("1100") && !c->set_access_code(std::string(65, '0')));
  }
  printf(bad ? "contract FAILED (%d)\n" : "contract ok\n", bad);
  return bad ? 1 : 0;
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s run|errors|contract ...\n", argv[0]); return 2; }
  try {
    if (!strcmp(argv[1], "errors")) return cmd_errors();
    if (!strcmp(argv[1], "contract")) return cmd_contract();
    if (!strcmp(argv[1], "run") && argc >= 6) return cmd_run(argc, argv);
  } catch (const std::exception& e) {
    fprintf(stderr, "exception: %s\n", e.what());
    return 3;
  }
  fprintf(stderr, "bad arguments\n");
  return 2;
}
