for lib in liborphx.so liborphx_seg16.so; do echo $lib; ORPHX_LIB=$PWD/orphics_b200/_lib/$lib python tools/bench_qe.py 8192 2 f64 TT 2>&1 | tail -1 | cut -c1-330; done
