"""Defaults shared across the package (values follow GPErks/constants.py:1-57)."""
import multiprocessing
import os
from pathlib import Path

from scipy.stats import norm

APPLICATION_NAME = "gperks_b200"

# logging
DEFAULT_LOG_FILE = (Path.home() / f"{APPLICATION_NAME}.log").as_posix()
DEFAULT_LOG_FORMAT = "%(levelname)s:%(asctime)s:%(module)s:%(funcName)s:L%(lineno)d: %(message)s"

# training
DEFAULT_TRAIN_MAX_EPOCH = 100
DEFAULT_TRAIN_SNAPSHOT_FREQUENCY = DEFAULT_TRAIN_MAX_EPOCH / 20
DEFAULT_TRAIN_SNAPSHOT_DIR = Path(os.getcwd(), "snapshot").as_posix()
DEFAULT_TRAIN_SNAPSHOT_SPLIT_TEMPLATE = "split_{split}"
DEFAULT_TRAIN_SNAPSHOT_RESTART_TEMPLATE = "restart_{restart}"
DEFAULT_TRAIN_SNAPSHOT_EPOCH_TEMPLATE = "epoch_{epoch}.pth"

# inference / cross validation
DEFAULT_INFERENCE_GRID_DIM = 50
DEFAULT_CROSS_VALIDATION_N_SPLITS = 5
DEFAULT_CROSS_VALIDATION_MAX_WORKERS = multiprocessing.cpu_count()

# global sensitivity analysis
DEFAULT_GSA_CONF_LEVEL = 0.95
DEFAULT_GSA_Z = norm.ppf(0.5 + DEFAULT_GSA_CONF_LEVEL / 2)
DEFAULT_GSA_N = 1024
DEFAULT_GSA_N_BOOTSTRAP = 100
DEFAULT_GSA_N_DRAWS = 1000
DEFAULT_GSA_SKIP_VALUES = None
DEFAULT_GSA_THRESHOLD = 0.01

# plot sizes (kept for API compatibility; plotting libraries are optional)
HEIGHT = 9.36111
WIDTH = 5.91667

# miscellanea
DEFAULT_EXP_N_RESTARTS = 3
DEFAULT_RANDOM_SEED = 8
DEFAULT_TMP_OUTFILE_DIR = Path(os.getcwd(), "tmp").as_posix()

# B200 path
DEFAULT_PREDICT_CHUNK_BYTES = 2 << 30      # HBM budget for one K* chunk in GPEmulator.predict

# precision="auto" in GPEmulator.predict / predict_device: the sliced-integer variance path (tcgen05 kind::i8, exact int32
# accumulation) is used when its int32 sums cannot overflow, the problem is large enough for the tensor-core tiles to pay, and
# the cancellation in  var = s + noise - ||L^-1 k*||^2  (amplification ~ outputscale / noise) leaves the error inside 1e-6;
# otherwise the FP64 DMMA path.  Candidates in order of speed (each is tried only inside its ratio limit):
#   N_pad <= 16384:  "int8x5w" five 8-bit slices (40 bits, 15 slice products)  for outputscale/noise <= 3e2
#                    "int8x6w" six  8-bit slices (48 bits, 21 products)        for outputscale/noise <= 1e5
#   N_pad <= 32768:  "int8x5w" as above, when the digits of the factorised L^-1 keep the int32 sums exact (8-bit digits are exact for
#                    ANY data only up to 16384 rows; above, 128 x the largest row sum of |digits| must stay below 2^31 --
#                    ExactGPEngine.i8_sums_exact / gpx_i8_digit_bound, measured 2 x below the limit at N_pad = 26240), then
#                    "int8x6"  six  7-bit slices (42 bits, 21 products)        for outputscale/noise <= 2e3
# Measured worst relative error of std against the FP64 path over points ON and 1e-4 beside training inputs (N = 2048 and 8192,
# profiles/r02_i8_modes_*.json) ~ c * outputscale/noise with c = 2.6e-10 (int8x5w), 1.2e-12 (int8x6w), 8e-11 (int8x6): the
# limits keep that product at or below 1e-7, a factor 10 inside the 1e-6 bar (tests/test_gpu_baseline_shapes.py measured
# 3.5e-7 for int8x5w at a ratio of 1000 on the N=4096, D=12 RBF emulator before the limit was lowered to 300).
AUTO_I8_SLICES = 6                  # K_hat^-1 for the gradient (gpx_lauum_i8): six 7-bit slices
AUTO_I8_MIN_N = 1024
AUTO_I8_MAX_N = 32768
AUTO_I8_WIDE_MAX_N = 16384
AUTO_I8_MAX_SCALE_TO_NOISE = 2.0e3
AUTO_I8_W5_MAX_SCALE_TO_NOISE = 3.0e2
AUTO_I8_W6_MAX_SCALE_TO_NOISE = 1.0e5
# ... and, once per factorisation, only if the sliced-integer std reproduces the FP64 (DMMA) std to this relative tolerance on
# the probe set: the 128 training inputs of smallest FP64 predictive variance (out of up to 4096 evenly strided ones) -- where
# s + noise - q cancels the most -- plus the same points moved by 1e-4 of the unit cube.  A candidate that fails hands over
# to the next more accurate one, then to "fp64".
AUTO_I8_PROBE_TOL = 2.5e-7
AUTO_I8_PROBE_POOL = 4096
AUTO_I8_PROBE_POINTS = 128

