// ORACLE / TEST INFRASTRUCTURE ONLY (see ref_capi.cc header).  Multi-threaded CPU timing
// harness over the reference's own classes; filled in by the bench milestone.
extern "C" int grref_bench_placeholder() { return 0; }
