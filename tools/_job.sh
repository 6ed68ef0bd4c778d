python tools/gemm_selftest.py --case dgrad_bnbwd_relu
python tools/gemm_selftest.py --case dgrad_bnbwd_lrelu_dense
python tools/gemm_selftest.py --case conv_fwd
