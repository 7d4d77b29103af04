# Second library with the run-time knock-out switches compiled in (SVX_FLAT_KNOCK=<bits>, timing experiments only):
#   SVX_LIB=$PWD/voxsrc2020_speaker_verification_b200/libsvx_knock.so SVX_FLAT_KNOCK=1 python tools/prof_step.py
set -e
cd "$(dirname "$0")/../voxsrc2020_speaker_verification_b200"
python -m voxsrc2020_speaker_verification_b200.build >/dev/null 2>&1 || (cd .. && python -m voxsrc2020_speaker_verification_b200.build >/dev/null)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr \
  -DSVX_KNOCK -c csrc/conv_flat.cu -o build/conv_flat_knock.o
objs=$(ls build/*.o | grep -v conv_flat)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libsvx_knock.so build/conv_flat_knock.o $objs -cudart static -Xlinker --no-undefined -ldl -lpthread -lrt
echo built libsvx_knock.so
