#!/bin/bash
# Runs every building-block experiment of tools/umma_probe.cu on the GPU box (one process each) and
# collects the verdict lines in gpurun_out/umma_probe.log.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=tools/umma_probe.bin
LOG=gpurun_out/umma_probe.log
: > $LOG
run() { echo "== $*" >> $LOG; timeout 60 $P "$@" >> $LOG 2>&1; echo "   exit=$?" >> $LOG; }
nvidia-smi --query-gpu=name,driver_version --format=csv >> $LOG 2>&1
run tmem
# TMA tile images: swizzle span, element size, box columns, box rows
run dump 128 2 64 128
run dump 128 1 128 256
run dump 64 1 64 128
run dump 32 2 16 64
run dump 0 8 8 32
run dump 128 8 16 8
run dump 0 8 128 8
# fp16 SS, both operands K-major SW128 through TMA (the vote kernel's B operand)
run gemm f16_ss_tma N=256 ksteps=4
run gemm f16_ss_tma_lbo0 N=256 ksteps=4 a_lbo=0 b_lbo=0
run gemm f16_ss_N64 N=64 ksteps=4
# A from tensor memory (scores), B K-major SW128 through TMA
run gemm f16_ts N=256 ksteps=4 a_src=2
# host-built images (thread-written layouts): K-major SW128
run gemm f16_ss_img N=256 ksteps=4 a_src=1 b_src=1
# K = 16 tiles with 32-byte rows: SW32 through TMA (score MMA: block columns of Xh), SW32 image, no-swizzle image
run gemm f16_k16_sw32_tma N=64 ksteps=1 a_swz=32 b_swz=32
run gemm f16_k16_sw32_img N=64 ksteps=1 a_swz=32 b_swz=32 a_src=1 b_src=1
run gemm f16_k16_none_img N=64 ksteps=1 a_swz=0 b_swz=0 a_src=1 b_src=1 a_lbo=128 a_sbo=256 b_lbo=128 b_sbo=256
run gemm f16_k16_amix N=64 ksteps=1 a_swz=0 b_swz=32 a_src=1 a_lbo=128 a_sbo=256
# int8: K-major SW128 both through TMA; signed and unsigned A
run gemm i8_ss_tma i8=1 N=256 ksteps=4
run gemm i8_ss_tma_u8 i8=1 N=256 ksteps=4 a_u8=1
# int8 with A MN-major SW128 (digits generated per row): LBO candidates
run gemm i8_amn_lbo16 i8=1 N=256 ksteps=4 a_u8=1 a_mn=1 a_src=1 a_lbo=16
run gemm i8_amn_lbo2048 i8=1 N=256 ksteps=4 a_u8=1 a_mn=1 a_src=1 a_lbo=2048
run gemm i8_amn_lbo1024 i8=1 N=256 ksteps=4 a_u8=1 a_mn=1 a_src=1 a_lbo=1024 a_sbo=1024
run gemm i8_amn_swap i8=1 N=256 ksteps=4 a_u8=1 a_mn=1 a_src=1 a_lbo=1024 a_sbo=2048
# fp32 accumulation behaviour of the tensor core over many instructions
run gemm f16_acc1 N=64 ksteps=4 real=1 repeat=1
run gemm f16_acc64 N=64 ksteps=4 real=1 repeat=64
run gemm f16_acc1024 N=64 ksteps=4 real=1 repeat=1024
run gemm f16_acc8192 N=64 ksteps=4 real=1 repeat=8192
cat $LOG
