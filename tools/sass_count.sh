# static SASS size of the solve kernels: bash tools/sass_count.sh [lib.so]
lib=${1:-fcc_qp_b200/libfccqp_b200.so}
tmp=$(mktemp -d); cp "$lib" $tmp/l.so; (cd $tmp && cuobjdump -xelf all l.so >/dev/null && for c in *.cubin; do nvdisasm -g -c $c > dis.txt; done)
python tools/sass_size.py $tmp/dis.txt fcc_qp_b200/csrc/fccqp_kernel.cuh
rm -rf $tmp
