/*
 * nearl_b200.h -- C ABI of libnearl_b200.so, the B200-native replacement for the hot path
 * behind Nearl's `all_actions` pybind11 module.
 *
 * Every entry point below takes plain pointers and sizes (no Python, no torch types) and
 * returns an int status (NRL_OK or an NRL_ERR_* code; the message is in
 * nrl_last_error()).  The first six are one-for-one twins of the functions the reference
 * exports from /root/reference/src/actions_py.cpp:297-351 and are what a maintainer
 * binds in place of that module.  Grids follow the reference's
 * index math: out[g] with x = g / dims[0] / dims[1], y = g / dims[0] % dims[1],
 * z = g % dims[0] (/root/reference/src/voxelize.cu:113-115,
 * /root/reference/src/marching_observers.cu:416-420), i.e. C order for the cubic grids
 * every reference configuration uses.
 *
 * Pointers are HOST pointers unless the function name ends in `_dev`.  Host entry points
 * are synchronous (results are in `out` on return).  `_dev` twins take device pointers
 * and a cudaStream_t (passed as void*; NULL = legacy default stream), enqueue all work
 * on that stream and return without synchronising.  All entry points of one device share one
 * workspace per lane; the library orders successive calls that arrive on different streams
 * with an event wait on the device (a call on stream B does not overwrite buffers the previous
 * call on stream A is still reading), so `_dev` calls may be issued from any streams, in any
 * interleaving, from one thread at a time per lane (calls are serialised by a mutex).
 */
#ifndef NEARL_B200_H
#define NEARL_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define NRL_OK 0
#define NRL_ERR_OBSERVABLE 1   /* "The observable type is not supported"   actions_py.cpp:133-140 */
#define NRL_ERR_AGGREGATION 2  /* "The aggregation type is not supported"  actions_py.cpp:141-148,207-214 */
#define NRL_ERR_FRAMES 3       /* more than 512 frames                     actions_py.cpp:151-153 */
#define NRL_ERR_SHAPE 4        /* inconsistent / non-positive sizes        actions_py.cpp:58-61,132 */
#define NRL_ERR_ARGUMENT 5     /* null pointer or invalid parameter */
#define NRL_ERR_NO_DEVICE 6    /* no usable CUDA device: there is NO CPU fallback */
#define NRL_ERR_PARSE 7        /* nrl_pdb_parse: the text is not plain enough for the column parser (the caller reads it line by line) */
#define NRL_ERR_CUDA 100       /* 100 + cudaError_t */

#define NRL_MAX_FRAME_NUMBER 512   /* /root/reference/src/constants.h:11 */

/* observable ids (/root/reference/src/constants.h:19-32) and reducer ids (:39-48) */
enum { NRL_EXISTENCE = 1, NRL_DIRECT_COUNT = 2, NRL_DISTINCT_COUNT = 3, NRL_MEAN_DISTANCE = 11,
       NRL_CUMULATIVE_WEIGHT = 12, NRL_DENSITY = 13, NRL_DISPERSION = 14, NRL_ECCENTRICITY = 15,
       NRL_RADIUS_OF_GYRATION = 16 };
enum { NRL_AGG_MEAN = 1, NRL_AGG_STD = 2, NRL_AGG_MEDIAN = 3, NRL_AGG_VARIANCE = 4, NRL_AGG_MAX = 5,
       NRL_AGG_MIN = 6, NRL_AGG_ENTROPY = 7, NRL_AGG_DRIFT = 8 };

/* ---- twins of the reference FFI ----------------------------------------------------- */

/* replaces all_actions.frame_voxelize  (do_voxelize, actions_py.cpp:45-89 -> voxelize_host,
 * voxelize.cu:292-337).  coords [n_atoms,3], weights [n_atoms], out [dims0*dims1*dims2].
 * auto_translate != 0 recentres `coords` IN PLACE first (cpuutils.h:340-357), as the
 * reference does. */
int nrl_frame_voxelize(float *coords, const float *weights, int n_atoms, const int *dims, float spacing,
                       float cutoff, float sigma, int auto_translate, float *out);

/* replaces all_actions.frame_observation (do_frame_observation, actions_py.cpp:268-294 ->
 * observe_frame_host, marching_observers.cu:519-553).  Like the reference it does not
 * validate type_obs: an unknown id yields zeros. */
int nrl_frame_observation(const float *coords, const float *weights, int n_atoms, const int *dims,
                          float spacing, float cutoff, int type_obs, float *out);

/* replaces all_actions.marching_observer (do_marching_observers, actions_py.cpp:112-166 ->
 * marching_observer_host, marching_observers.cu:435-516).  coords [n_frames,n_atoms,3],
 * weights [n_frames*n_atoms]. */
int nrl_marching_observer(const float *coords, const float *weights, int n_frames, int n_atoms,
                          const int *dims, float spacing, float cutoff, int type_obs, int type_agg,
                          float *out);

/* replaces all_actions.density_flow (do_traj_voxelize, actions_py.cpp:188-229 ->
 * trajectory_voxelization_host, voxelize.cu:350-415).  Frames beyond 512 are ignored by
 * the reducer, as in the reference (voxelize.cu:403). */
int nrl_density_flow(const float *coords, const float *weights, int n_frames, int n_atoms, const int *dims,
                     float spacing, float cutoff, float sigma, int type_agg, float *out);

/* replaces all_actions.aggregate (do_aggregation, actions_py.cpp:232-252 -> aggregate_host,
 * gpuutils.cu:80-105).  series [n_frames, n_voxels] frame-major.  An unknown type_agg
 * yields zeros (gpuutils.cu:73-76). */
int nrl_aggregate(const float *series, int n_frames, long long n_voxels, int type_agg, float *out);

/* replaces all_actions.summation (do_summation, actions_py.cpp:254-266 -> sum_reduction_host,
 * gpuutils.cu:107-134). */
int nrl_summation(const float *arr, long long n, float *out);

/* ---- device-pointer twins (stream-ordered, no synchronisation) ------------------------ */
int nrl_frame_voxelize_dev(const float *coords, const float *weights, int n_atoms, const int *dims,
                           float spacing, float cutoff, float sigma, float *out, void *stream);
int nrl_frame_observation_dev(const float *coords, const float *weights, int n_atoms, const int *dims,
                              float spacing, float cutoff, int type_obs, float *out, void *stream);
int nrl_marching_observer_dev(const float *coords, const float *weights, int n_frames, int n_atoms,
                              const int *dims, float spacing, float cutoff, int type_obs, int type_agg,
                              float *out, void *stream);
int nrl_density_flow_dev(const float *coords, const float *weights, int n_frames, int n_atoms,
                         const int *dims, float spacing, float cutoff, float sigma, int type_agg, float *out,
                         void *stream);
int nrl_aggregate_dev(const float *series, int n_frames, long long n_voxels, int type_agg, float *out,
                      void *stream);

/* ---- batched featurisation: one call per trajectory ----------------------------------- */
/* A crop is one focus box on one window of trajectory frames; it restates Feature.query /
 * crop / DynamicFeature.query (/root/reference/nearl/features.py:92-123,426-483,1329-1368)
 * on the device.  x' = ((x - focal) + center) - spacing/2 evaluated in the trajectory's
 * dtype; an atom is kept iff lo < x' < hi on every axis, then by-residue expansion,
 * selection mask and the MAX_ALLOWED_ATOMS cap apply. */
typedef struct {
  int frame_start;       /* first trajectory frame of the window */
  int n_frames;          /* frames in the window (1 for static features) */
  float focal[3];        /* focal point, float32 as stored in Featurizer.FOCALPOINTS */
  float center[3];       /* Feature.center */
  float half_spacing_f32;
  double half_spacing_f64;
  double lo[3], hi[3];
  int byres;
  int selection;         /* row of `selections`, -1 = every atom */
  int max_atoms;         /* 0 = no cap */
  int sort_table;        /* row of `value_orders` to emit atoms in (distinct_count), -1 = topology order */
  /* Featurizer.loop_by_residue (/root/reference/nearl/featurizer.py:471-554) featurises a SLICE of the
   * topology per task: only atoms atom_begin <= a < atom_end are looked at (atom_end <= 0: every atom).
   * Per-atom tables (weights, selections, residue index) are read at a + table_shift: the reference re-caches
   * a feature only when the slice length changes (features.py:452-457), so a slice can be read with the
   * tables of an earlier slice of the same length. */
  int atom_begin, atom_end;
  int table_shift;
  double focal_f64[3];   /* focal point for float64 coordinates (the float32 focal point widened, or the float64
                            mean of a residue slice, featurizer.py:503) */
} nrl_crop_t;

/* A task is one (feature, crop): mode 0 = Gaussian voxelisation, otherwise an observable id;
 * agg 0 = the crop's single frame as is (static features), otherwise a reducer id. */
typedef struct {
  int crop;
  int mode;
  int agg;
  int weight_table;      /* row of `weight_tables` */
  int dims[3];
  float spacing, cutoff, sigma;
  int rank_table;        /* distinct_count only: row of `weight_tables` holding, as floats, the rank (0..31) of every
                            atom's value among the distinct values of `weight_table` (-1 for the 99999.0 placeholder,
                            which is never counted); -1 = no ranks: the crop must then be value ordered (sort_table) */
  /* One task split over several GPUs: only the slab x_begin <= ix < x_begin + x_count of the grid's slowest axis is
   * computed (focus box, crop and normalisation are those of the whole grid, so the slab has the bits of the same rows
   * of an unsplit task) and the task's result is that slab alone, x_count * dims[1] * dims[2] floats.  x_count 0 = the
   * whole grid.  x_begin must be a multiple of 8 (the tile edge), x_count too unless the slab ends with the grid;
   * cubic grids only. */
  int x_begin, x_count;
} nrl_task_t;

typedef struct {
  const void *xyz;               /* [n_frames_total, n_atoms, 3], host (pinned for speed) or device */
  int xyz_is_f64;
  int xyz_on_device;
  int n_frames_total;
  int n_atoms;
  const float *weight_tables;    /* [n_weight_tables, n_atoms] host */
  int n_weight_tables;
  const unsigned char *selections; /* [n_selections, n_atoms] host, may be NULL */
  int n_selections;
  const int *residue_index;      /* [n_atoms] dense residue index, may be NULL when no crop is by-residue */
  int n_residues;
  const int *value_orders;       /* [n_weight_tables, n_atoms] argsort of each weight table, may be NULL */
  const nrl_crop_t *crops;
  int n_crops;
  const nrl_task_t *tasks;
  int n_tasks;
  float *out;                    /* host (pinned for speed) or device: task results back to back, task order */
  int out_on_device;
  int *atom_counts;              /* optional host [total crop frames]: atoms kept per crop frame */
  int flags;                     /* NRL_BATCH_* */
  int lane;                      /* 0 .. NRL_MAX_LANES-1: independent stream + workspace (see nrl_lane_synchronize) */
} nrl_batch_t;

#define NRL_BATCH_NO_SYNC 1      /* enqueue only: results are valid after nrl_lane_synchronize(lane); host buffers
                                    (xyz, out) must be page-locked and stay alive until then */
#define NRL_MAX_LANES 8

int nrl_featurize_batch(const nrl_batch_t *batch);

/* *out = 1 when a grid of this geometry gets a cell list (rank-based distinct_count needs one), else 0 */
int nrl_cell_list_fits(const int *dims, float spacing, float cutoff, int *out);

/* ---- trajectory preparation ---------------------------------------------------------------- */
/* Least-squares superposition of every frame onto frame `ref_frame` over the atoms `fit_index`, IN PLACE
 * (what MisatoTraj(..., superpose=True) asks pytraj for: pt.superpose(traj, mask="@CA"),
 * /root/reference/nearl/io/traj.py:398-404).  xyz [n_frames, n_atoms, 3] float32 or float64, host or device
 * (on_device: `stream`-ordered, no synchronisation); rmsd (optional, n_frames floats in the same memory space)
 * receives the fitted RMSD of every frame.  The reference frame is left untouched. */
int nrl_superpose(void *xyz, int xyz_is_f64, int n_frames, int n_atoms, const int *fit_index, int n_fit, int ref_frame,
                  int on_device, float *rmsd, void *stream);

/* Focal points of Featurizer.parse_focus (/root/reference/nearl/featurizer.py:357-372): out[f] = centre of geometry of the
 * atoms `sel` in trajectory frame frames[f], float32, with the arithmetic of numpy's np.mean(frame[sel], axis=0) (rows added
 * in order in the trajectory's precision, divided in double).  Plain host code: xyz [n_frames_total, n_atoms, 3] in host
 * memory, float32 or float64. */
int nrl_focal_means(const void *xyz, int xyz_is_f64, int n_frames_total, int n_atoms, const int *frames, int n_frames,
                    const int *sel, int n_sel, float *out);

/* Ingestion (host code, no device): the ATOM / HETATM records of a PDB text -- what pt.load hands
 * Trajectory for "complex.pdb" (/root/reference/nearl/io/traj.py:97-106).  Two calls: with xyz == NULL and columns == NULL the
 * counts only (records in the file, records of the first model, models), then the fill: xyz [n_records, 3] float64 in file
 * order, columns [n_first, 16] bytes per record of the first model: atom name (PDB columns 13-16), residue name (18-21),
 * chain (22), residue number + insertion code (23-27), element (77-78, blank when the line ends before them).  A frame ends at
 * a line that starts with END (ENDMDL included) and has records in front of it.  NRL_ERR_PARSE: the text is not plain enough
 * (non-ASCII bytes, a bare carriage return, a record shorter than 54 columns, a coordinate field that is not
 * [blank][sign]digits[.digits][blank], models of different sizes, no record): the caller parses it line by line. */
int nrl_pdb_parse(const char *text, long long n_bytes, long long *n_records, long long *n_first, long long *n_frames, double *xyz,
                  char *columns);

/* ---- utilities -------------------------------------------------------------------------- */
/* buff_dim exactly as the device evaluates (cutoff+spacing)/spacing (voxelize.cu:50, div.approx) */
int nrl_buff_dim(float cutoff, float spacing, int *out);
int nrl_device_count(int *out);
int nrl_set_device(int device);
int nrl_synchronize(void);
/* wait for everything enqueued on `lane` (batches submitted with NRL_BATCH_NO_SYNC); batches on
 * different lanes overlap, e.g. the upload of one trajectory under the kernels of the previous one */
int nrl_lane_synchronize(int lane);
/* device-side joins between a caller's stream and a lane: nrl_lane_fork makes the lane's future work start after what
 * `stream` holds now, nrl_lane_join makes `stream` wait for what the lane holds now (no host synchronisation) */
int nrl_lane_fork(int lane, void *stream);
int nrl_lane_join(int lane, void *stream);
/* enable != 0: enqueue all work of the host entry points and of nrl_featurize_batch on `stream`
 * (a cudaStream_t, e.g. torch's current stream) instead of the library's own stream */
int nrl_set_stream(void *stream, int enable);
/* tuning / test switches.  "force_scan" != 0: never use the cell-list kernels (every task runs on the
 * full-frame scan kernels; results are identical up to float summation order).  "no_fuse" != 0: observers of
 * one focus box do not share a sweep across observables (results are bit-identical either way) */
/* "no_fork" != 0: the per-mode launches of a batch are not forked onto side streams (they run one after the other: per-kernel
 * event times are then those of the kernels alone); "crop_legacy" != 0: whole-frame crops never take the streaming kernel; "crop_stream_min" = n: the shortest topology
 * slice that does (default 2048; the forms select the same atoms in the same order) */
int nrl_set_option(const char *name, int value);
/* per-kernel device timing: CUDA events around every launch of this library, accumulated per
 * category, on every lane of the current device; nrl_profile_read synchronises the lanes and copies
 * NRL_PROFILE_CATEGORIES totals (ms, launches) summed over them */
#define NRL_PROFILE_CATEGORIES 19
int nrl_profile(int enable);
int nrl_profile_read(double *ms, long long *counts, int reset);
const char *nrl_profile_name(int category);
/* page-lock / unlock caller memory so that the copies inside the host entry points run as DMA */
int nrl_host_register(void *ptr, long long bytes);
int nrl_host_unregister(void *ptr);
/* kernels launched by this library since load (bench.py reports the delta as gpu_launches) */
long long nrl_kernel_launches(void);
const char *nrl_last_error(void);
const char *nrl_version(void);

#ifdef __cplusplus
}
#endif
#endif
