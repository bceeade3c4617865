# per-CTA cycle accounting of the fused four-step kernel (FFB_L2_DEBUG) for a few scheduler settings; usage: bash tools/l2four_debug.sh [shape] [dtype]
S=${1:-8192x8192}; T=${2:-f64}
for v in "FFB_L2_CHUNK=2" "FFB_L2_CHUNK=1" "FFB_L2_CHUNK=4" "FFB_L2_AHEAD=10" "FFB_L2_AHEAD=15" "FFB_L2_PF=0" "FFB_L2_ACQ=0"; do echo "--- $v"; env $v FFB_L2_DEBUG=1 python tools/run_fft_once.py $S $T 2 2>&1 | tail -3 | head -2; done
