/* hpxfft_b200.h -- C ABI of libhpxfft_b200.so
 *
 * B200-native (sm_100a) replacement for the one hot path of constracktor/HPX-FFT: the FP64 2-D
 * real-to-complex FFT behind hpxfft::shared::loop and hpxfft::distributed::loop.
 * Every entry point cites the reference interface it replaces (paths relative to the reference
 * repository root).  The reference has no FFI of its own -- its boundary is the C++ class surface
 * (core/include/hpxfft/shared/loop.hpp:15-31, core/include/hpxfft/distributed/loop.hpp:17-31); the
 * drop-in classes in include/hpxfft/ and the ctypes mirror in hpx-fft_b200/ are thin wrappers over
 * exactly these functions.
 *
 * Process model: SPMD, one process ("locality") per GPU, like the reference's
 * hpx.run_hpx_main!=1 launch (examples/hpxfft/distributed_loop_2d.cpp:137-142).  The host runtime
 * (HPX, MPI, torch.distributed, ...) only has to broadcast/all-gather a few hundred bytes at plan
 * creation; all data movement afterwards is NCCL or peer-to-peer stores over NVLink.
 *
 * Data contract (core/include/hpxfft/util/vector_2d.hpp:12-16,198-213): row-major doubles,
 * n_x_local rows x n_col columns, n_col = 2*(ny/2+1); input occupies columns [0, ny), output is the
 * interleaved (re,im) Hermitian half  Z[kx][ky], ky = 0..ny/2  in the same rows.  Forward,
 * unnormalised (core/src/util/adapter_fftw.cpp:9,28-29).
 *
 * All functions return 0 on success or a negative HPXFFT_B200_E* code; hpxfft_b200_last_error()
 * returns a thread-local description.  There is no CPU fallback: without a CUDA device every
 * compute entry point fails with HPXFFT_B200_ECUDA.
 */
#ifndef HPXFFT_B200_H_INCLUDED
#define HPXFFT_B200_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HPXFFT_B200_VERSION 200

#define HPXFFT_B200_OK 0
#define HPXFFT_B200_EINVAL (-1)     /* bad argument / unsupported size                          */
#define HPXFFT_B200_EPLANFLAG (-2)  /* unknown plan flag  -> std::invalid_argument in the wrapper */
#define HPXFFT_B200_ECOMMFLAG (-3)  /* unknown comm flag  -> message, no transform (loop.cpp:342) */
#define HPXFFT_B200_ECUDA (-4)      /* CUDA runtime / driver error                               */
#define HPXFFT_B200_ENCCL (-5)      /* NCCL error                                                */
#define HPXFFT_B200_ESTATE (-6)     /* call sequence error (e.g. p2p plan executed before import) */

#define HPXFFT_B200_UNIQUE_ID_BYTES 128 /* == sizeof(ncclUniqueId) */
#define HPXFFT_B200_IPC_HANDLE_BYTES 64 /* == sizeof(cudaIpcMemHandle_t) */

/* synthetic input patterns for hpxfft_b200_fill (definitions shared with oracle/oracle.py) */
#define HPXFFT_B200_PATTERN_RAMP 0      /* v(i,j) = j   examples/hpxfft/shared_loop_2d.cpp:33-40 */
#define HPXFFT_B200_PATTERN_UNIFORM 1   /* v(i,j) = u(splitmix64(seed ^ (i_global*ny + j))) in [-1,1) */
#define HPXFFT_B200_PATTERN_SEPARABLE 2 /* v(i,j) = sum_{r<4} a_r(i_global) b_r(j)            */

typedef struct hpxfft_b200_plan hpxfft_b200_plan; /* opaque: device buffers, streams, events, comm */

int hpxfft_b200_version(void);
const char *hpxfft_b200_last_error(void);
/* number of visible CUDA devices (0 without a GPU; never fails) */
int hpxfft_b200_device_count(void);

/* Communicator bootstrap.  Replaces hpx::collectives::create_communicator(basename, num_sites,
 * this_site) (core/src/distributed/loop.cpp:324-327,337-340): rank 0 obtains an id, the host
 * runtime broadcasts the 128 bytes, every rank passes them to hpxfft_b200_create. */
int hpxfft_b200_get_unique_id(void *id_out /* HPXFFT_B200_UNIQUE_ID_BYTES */);

/* Column ownership after exchange #1: rank q owns ky in [c0, c0 + w), c0 = q*floor(cy/nranks), the
 * last rank absorbs cy mod nranks (cy = ny/2+1 is odd for power-of-two ny, so the split is always
 * uneven).  The reference uses n_y_local = dim_c_y/L and silently drops the remainder
 * (core/src/distributed/loop.cpp:288-289).  Pure host arithmetic, no GPU needed. */
int hpxfft_b200_partition(size_t cy, int nranks, int rank, size_t *c0, size_t *w);

/* Replaces shared::loop::initialize (core/src/shared/loop.cpp:158-189) and
 * distributed::loop::initialize (core/src/distributed/loop.cpp:275-347): dimension inference
 * (dim_c_y = n_col/2, ny = 2*dim_c_y-2, nx = n_x_local*nranks), buffer allocation, "planning"
 * (twiddle tables, kernel selection) and communicator set-up.
 *   n_x_local : rows of this locality's slab (vector_2d::n_row)
 *   n_col     : doubles per row (vector_2d::n_col), even, >= 4
 *               supported sizes: ny = n_col-2 with ny/2 a power of two <= 65536 (ny <= 131072) and
 *               nx = n_x_local*nranks a power of two <= 2^18 take the power-of-two kernels; lengths
 *               t * 2^a with a small odd t (rows: t < 32 and ny/2 <= 8192; columns: t <= 127) the mixed-radix
 *               kernels; any other even ny / any nx up to 131072 Bluestein's chirp-z (<= 8192: direct DFT);
 *               anything else is rejected with HPXFFT_B200_EINVAL
 *   rank, nranks : this locality / number of localities (hpx::get_locality_id / get_num_localities)
 *   device    : CUDA device ordinal for this rank, or -1 for the current device
 *   comm_flag : NULL (shared::loop, nranks must be 1) | "scatter" | "all_to_all"   (reference modes,
 *               distributed/loop.cpp:156-179) | "p2p" (fused peer-store variant, extension).
 *               "scatter"    = P rooted scatters, one NCCL send/recv group per root, issued in root order
 *                              (scatter_to / scatter_from on P communicators, distributed/loop.cpp:39-69,158-167);
 *               "all_to_all" = one personalised all-to-all per exchange (distributed/loop.cpp:72-84).  Its
 *                              transport is chosen by the environment variable HPXFFT_B200_A2A:
 *                              "fused"         the FFT kernels store straight into the owners' windows over NVLink: the
 *                                              exchange overlaps the producing kernel tile by tile, no staging buffer, no
 *                                              unpack pass (default while a slab row is at most 512 KB, i.e. ny <= 65536);
 *                              "ce"            copy-engine peer copies over the IPC windows, cut into sub-slab chunks that
 *                                              overlap the FFT kernels, landing exchange #2 directly in the destination slab
 *                                              (default for longer rows, e.g. 131072^2);
 *                              "nccl"          one grouped ncclSend/ncclRecv exchange + unpack kernel;
 *               "p2p"        = the FFT kernels store straight into the owners' windows over NVLink.
 *               Plans whose hpxfft_b200_ipc_count() is non-zero need the export / all-gather / import
 *               handshake below before the first execute.
 *   plan_flag : "estimate" | "measure" | "patient" | "exhaustive"  (util/adapter_fftw.hpp:22-44)
 *   unique_id : HPXFFT_B200_UNIQUE_ID_BYTES from rank 0's hpxfft_b200_get_unique_id; may be NULL
 *               when nranks == 1 */
int hpxfft_b200_create(hpxfft_b200_plan **out, size_t n_x_local, size_t n_col, int rank, int nranks,
                       int device, const char *comm_flag, const char *plan_flag, const void *unique_id);

/* Peer windows (copy-engine and fused transports): export this rank's receive-window handles
 * (count * HPXFFT_B200_IPC_HANDLE_BYTES, count = hpxfft_b200_ipc_count(); 0 = nothing to do), all-gather
 * them in rank order, import on every rank.  Replaces the per-exchange buffers HPX serialises
 * (core/src/distributed/loop.cpp:51,75). */
int hpxfft_b200_ipc_count(const hpxfft_b200_plan *);
int hpxfft_b200_ipc_export(hpxfft_b200_plan *, void *handles_out);
int hpxfft_b200_ipc_import(hpxfft_b200_plan *, const void *all_handles /* nranks * count * 64 B */);

/* name of the exchange transport the plan settled on: "none", "nccl", "nccl-rooted", "nccl-pipelined",
 * "copy-engine", "fused-peer-store" */
const char *hpxfft_b200_transport(const hpxfft_b200_plan *);
/* Pins the calling thread (and what it allocates by first touch, e.g. the page-locked vector_2d storage) to
 * the CPUs of the NUMA node the device hangs off, so that host<->device copies of all ranks of a node do
 * not funnel through one socket.  device = -1: current device. */
int hpxfft_b200_bind_host_to_device(int device);

/* Host <-> device staging of the slab (the by-value vector_2d hand-over of initialize /
 * "return std::move(values_vec_)", core/src/shared/loop.cpp:161,112). */
int hpxfft_b200_upload(hpxfft_b200_plan *, const double *host_slab);
int hpxfft_b200_download(hpxfft_b200_plan *, double *host_slab);
/* host_slab may also be a DEVICE pointer (unified addressing): the copy is then device-to-device and the
 * caller keeps its data on the GPU (initialize from device memory, no PCIe).
 * download_tile copies the sub-block rows [row0, row0+nrows) x doubles [col0, col0+ncols) into a dense host
 * array -- sampled checks of slabs that are too large to bring back whole. */
int hpxfft_b200_download_tile(hpxfft_b200_plan *, size_t row0, size_t nrows, size_t col0, size_t ncols, double *host_out);
/* on-device synthetic input (no PCIe); row index is global: rank*n_x_local + i */
int hpxfft_b200_fill(hpxfft_b200_plan *, int pattern, uint64_t seed);

/* Replaces shared::loop::fft_2d_r2c_par / _seq (core/src/shared/loop.cpp:56-155) and
 * distributed::loop::fft_2d_r2c (core/src/distributed/loop.cpp:130-272).  Device-resident,
 * asynchronous launch + stream synchronise; records the reference's timer keys from CUDA events.
 * Collective across ranks when nranks > 1. */
int hpxfft_b200_execute(hpxfft_b200_plan *);
/* Back-to-back operation: execute_async only enqueues; synchronize waits and sets every
 * measurement key to the AVERAGE over the transforms enqueued since the last reset_timers / execute
 * (at most the 64 most recent; key "timer_samples" says how many). */
int hpxfft_b200_execute_async(hpxfft_b200_plan *);
int hpxfft_b200_synchronize(hpxfft_b200_plan *);
int hpxfft_b200_reset_timers(hpxfft_b200_plan *);
/* upload + execute + download in one call: what initialize()+fft_2d_r2c() cost end to end */
int hpxfft_b200_transform(hpxfft_b200_plan *, double *host_slab_inout);
/* the same three steps only ENQUEUED on the plan's stream (host_slab_inout must stay valid and should be
 * page-locked); finish with hpxfft_b200_synchronize.  Two plans driven alternately keep PCIe busy in both
 * directions: the download of transform i overlaps the upload of transform i+1. */
int hpxfft_b200_transform_async(hpxfft_b200_plan *, double *host_slab_inout);

/* Building blocks of a fully asynchronous round trip (the agas client surfaces, core/include/hpxfft/shared/agas.hpp:19-24,
 * core/include/hpxfft/distributed/agas.hpp:21-26, return futures): upload_async / execute_async / download_async only
 * ENQUEUE on the plan's stream; on_complete enqueues a host function (cudaLaunchHostFunc) that runs when everything
 * enqueued before it has finished -- it fulfils the future without a thread blocked on the GPU.  The callback must not
 * call back into this library or CUDA.  Host buffers should be page-locked (hpxfft_b200_host_alloc) and must stay
 * valid until the callback has run. */
typedef void (*hpxfft_b200_callback)(void *user);
int hpxfft_b200_upload_async(hpxfft_b200_plan *, const double *host_slab);
int hpxfft_b200_download_async(hpxfft_b200_plan *, double *host_slab);
int hpxfft_b200_on_complete(hpxfft_b200_plan *, hpxfft_b200_callback fn, void *user);

/* Runs exchange #1 or #2 (which = 1 | 2) ALONE `reps` times on whatever the staging buffers hold and
 * returns the average milliseconds -- the NVLink roofline of the transport without the FFT kernels.
 * Collective; the slab contents are undefined afterwards.  HPXFFT_B200_ESTATE for the fused transport. */
int hpxfft_b200_bench_exchange(hpxfft_b200_plan *, int which, int reps, double *ms_out);

/* Replaces loop::get_measurement (core/src/shared/loop.cpp:192, distributed/loop.cpp:350): seconds;
 * keys total, first_fftw, first_trans, second_fftw, second_trans, plan, plan_flops (+ first_split,
 * first_comm, second_split, second_comm for distributed; extensions h2d, d2h, rows_kernel,
 * cols_kernel, cols_levelA_kernel, cols_levelB_kernel, first_comm_span, second_comm_span, timer_samples).  Unknown key -> 0.0 like std::map::operator[]. */
double hpxfft_b200_measurement(const hpxfft_b200_plan *, const char *key);

/* Replaces loop::write_plans_to_file (core/src/shared/loop.cpp:194-212): appends a text description
 * of the row and column kernel plans.  HPXFFT_B200_EINVAL if the file cannot be opened. */
int hpxfft_b200_write_plans(const hpxfft_b200_plan *, const char *file_path);

/* zero-copy access for harnesses: device slab pointer (n_x_local x n_col doubles) and the stream
 * all work of this plan is ordered on (a cudaStream_t). */
void *hpxfft_b200_device_ptr(hpxfft_b200_plan *);
void *hpxfft_b200_stream(hpxfft_b200_plan *);
/* number of this library's kernels launched by one hpxfft_b200_execute */
int hpxfft_b200_launches_per_execute(const hpxfft_b200_plan *);

void hpxfft_b200_destroy(hpxfft_b200_plan *);

/* Page-locked host storage for vector_2d::values_ (the reference uses `new T[size_]`,
 * core/include/hpxfft/util/vector_2d.hpp:97-110): lets upload/download/transform run at full PCIe
 * speed.  Returns NULL on failure (see hpxfft_b200_last_error). */
void *hpxfft_b200_host_alloc(size_t bytes);
void hpxfft_b200_host_free(void *ptr);

/* Backend seam of util::fftw_adapter (core/src/util/adapter_fftw.cpp:12-15, 32-35), exposed for
 * kernel-level parity tests.  Host buffers, staged through the device.
 *   r2c_rows : `batch` padded rows of n_col = 2*(ny/2+1) doubles, in place  (r2c_1d::execute)
 *   c2c_cols : forward c2c along the FIRST axis of a row-major [n][width] complex array, in place
 *              (c2c_1d::execute on every column; the reference runs it on rows of the transposed
 *              array, here the transposes are fused away) */
int hpxfft_b200_r2c_rows(double *host_rows, size_t batch, size_t n_col, int device);
int hpxfft_b200_c2c_cols(double *host_data, size_t n, size_t width, int device);
/* variant 0 = the kernels a plan of this length launches (the persistent fused four-step kernel for n > 256),
 * variant 1 = the unfused launches (single tile / level A + level B) */
int hpxfft_b200_c2c_cols_variant(double *host_data, size_t n, size_t width, int device, int variant);

#ifdef __cplusplus
}
#endif
#endif /* HPXFFT_B200_H_INCLUDED */
