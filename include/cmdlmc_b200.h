/*
 * cmdlmc_b200.h -- C ABI of the B200-native cMD/LMC hot path (libcmdlmc_b200.so).
 *
 * The reference has no FFI: its only native seam is the Cython `cdef class AtomBox` method
 * table (mdlmc/cython_exts/LMC/PBCHelper.pxd:1-43) plus the duck-typed Python protocols that
 * main.py:73-158 wires together.  Each entry point below names the reference interface it
 * replaces (paths relative to the reference root).  INTEGRATION.md shows the ctypes binding a
 * reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative CMD_E* code on failure; the message is
 *     available from cmd_last_error() (thread-local).  Nothing throws across the boundary.
 *   - plain pointers and sizes only.  `h_` arguments are HOST pointers borrowed for the call,
 *     `d_` arguments are DEVICE pointers (current device) used on the stream set with
 *     cmd_set_stream().  Positions are float64 [n][3], C-contiguous, like the reference
 *     (PBCHelper.pyx:60-61,78-79,88).
 *   - one host thread per process drives one GPU (one process per GPU); not re-entrant.
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with
 *     CMD_ENODEV.
 */
#ifndef CMDLMC_B200_H
#define CMDLMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMD_ABI_VERSION 1

#define CMD_OK 0
#define CMD_EINVAL (-1)    /* bad argument */
#define CMD_ENODEV (-2)    /* no CUDA device / cmd_init not called */
#define CMD_ECUDA (-3)     /* CUDA runtime error (see cmd_last_error) */
#define CMD_ENOMEM (-4)    /* allocation failed */
#define CMD_ECAPACITY (-5) /* an internal fixed-capacity buffer overflowed; retry with more */
#define CMD_ESTATE (-6)    /* object used in the wrong state */

/* jump-rate function kinds (mdlmc/LMC/jumprate_generators.py:14-43; legacy kinds from the
 * specification text mdlmc/IO/config_parser.py:322-349, parity unpinned) */
#define CMD_RATE_FERMI 0       /* par = a, b, c            w = a / (1 + exp((x - b) / c))      */
#define CMD_RATE_FERMI_ANGLE 1 /* par = a, b, c, theta0    0 where theta < theta0 else Fermi   */
#define CMD_RATE_AE 2          /* par = A, a, b, d0, T     w = A exp(-E/(kB T)), E = a u / sqrt(b + 1/u^2), u = x - d0 */
#define CMD_RATE_EXP 3         /* par = a, b               w = a exp(b x)                      */
#define CMD_RATE_NPAR 8

/* distance conversions of AtomBoxWater* (PBCHelper.pyx:306-351) */
#define CMD_CONV_NONE 0
#define CMD_CONV_LINEAR 1 /* par = a, b, -, left, right */
#define CMD_CONV_RAMP 2   /* par = a, b, d0, left, right */

typedef struct cmd_box cmd_box;
typedef struct cmd_topo cmd_topo;
typedef struct cmd_kmc cmd_kmc;

/* ---------------------------------------------------------------- lifecycle ------------- */
int cmd_abi_version(void);
const char *cmd_last_error(void);
/* Selects the CUDA device of this process (one process per GPU) and creates the library's
 * stream-ordered scratch state.  Fails with CMD_ENODEV when no device is present. */
int cmd_init(int device);
int cmd_shutdown(void);
int cmd_device_count(int *n);
/* Stream used by every kernel the library launches (a cudaStream_t passed as void*; NULL
 * selects the legacy default stream).  PyTorch callers pass
 * torch.cuda.current_stream().cuda_stream so that torch.cuda.Event timing sees the work. */
int cmd_set_stream(void *cuda_stream);
int cmd_sync(void);
/* Number of kernels this library has launched since cmd_init (bench.py's gpu_launches). */
int64_t cmd_launch_count(void);
/* Measured-peak helper: runs a dependent-free DFMA loop on every SM and returns the achieved
 * FP64 rate in TFLOP/s (the roofline denominator of the FP64-bound kernels). */
int cmd_fp64_peak(int iters, double *tflops);
/* The same for shared memory: conflict-free 16-byte loads on every SM, GB/s over the whole GPU
 * (the roofline denominator of the KMC replica kernel, which streams the pair lists out of a
 * shared-memory ring). */
int cmd_smem_peak(int iters, double *gbs);

/* ---------------------------------------------------------------- AtomBox --------------- */
/* Replaces AtomBoxCubic.__cinit__ (n_values == 3, PBCHelper.pyx:216-226) and
 * AtomBoxMonoclinic.__cinit__ (n_values == 9, rows = cell vectors, PBCHelper.pyx:248-260). */
int cmd_box_create(const double *h_periodic_boundaries, int n_values,
                   const int h_box_multiplier[3], cmd_box **out);
/* The reference inverts h with np.linalg.inv (PBCHelper.pyx:259); a caller that wants device
 * results bit-identical with the reference's h_inv hands that matrix over here. */
int cmd_box_set_hinv(cmd_box *box, const double h_hinv[9]);
/* number of periodic images the fast pair filter has to look at for this cell (0 for ortho) */
int cmd_box_n_images(const cmd_box *box);
/* AtomBoxWaterLinearConversion / AtomBoxWaterRampConversion (PBCHelper.pyx:306-351). */
int cmd_box_set_conversion(cmd_box *box, int conv_kind, const double h_par[5]);
/* public attributes of the Cython class (PBCHelper.pxd:3-8,40-43): periodic_boundaries_extended
 * (3 or 9 values), pbc_matrix[9], h[9], h_inv[9]; any output pointer may be NULL. */
int cmd_box_query(const cmd_box *box, double *h_pbc_extended, double *h_pbc_matrix, double *h_h,
                  double *h_hinv);
void cmd_box_destroy(cmd_box *box);

/* AtomBox.length (PBCHelper.pyx:74-85): out[i] = |min-image(b[i] - a[i])| (27-image search for
 * general cells, numpyatom.pyx:101-123; water conversion applied). */
int cmd_length(const cmd_box *box, const double *h_a, const double *h_b, int64_t n, double *h_out);
/* AtomBox.distance (PBCHelper.pyx:56-70): out[i][3] = wrapped vector b[i] - a[i] (fractional
 * wrap only for general cells, numpyatom.pyx:61-74). */
int cmd_distance(const cmd_box *box, const double *h_a, const double *h_b, int64_t n,
                 double *h_out);
/* AtomBox.length_all_to_all (PBCHelper.pyx:88-95): out[i][j] = length(a[i], b[j]). */
int cmd_length_all_to_all(const cmd_box *box, const double *h_a, int64_t n, const double *h_b,
                          int64_t m, double *h_out);
/* AtomBox.angle (PBCHelper.pyx:133-134,237-239,273-275): angle at a2 between a1 and a3. */
int cmd_angle(const cmd_box *box, const double *h_a1, const double *h_a2, const double *h_a3,
              int64_t n, double *h_out);
/* AtomBox.next_neighbor (PBCHelper.pyx:153-167): first index of the smallest length. */
int cmd_next_neighbor(const cmd_box *box, const double *h_pos, const double *h_frame, int64_t n,
                      int *idx, double *dist);
/* AtomBox.position_extended_box (PBCHelper.pyx:34-53), host arithmetic. */
int cmd_position_extended_box(const cmd_box *box, int index, const double *h_frame, int n_atoms,
                              double h_out[3]);
/* AtomBox.next_neighbor_extended_box (PBCHelper.pyx:169-185). */
int cmd_next_neighbor_extended_box(const cmd_box *box, int index_1, const double *h_frame_1,
                                   int n1, const double *h_frame_2, int n2, int *idx,
                                   double *dist);

/* Device-pointer forms of the same operations (inputs already resident in HBM). */
int cmd_length_dev(const cmd_box *box, const double *d_a, const double *d_b, int64_t n,
                   double *d_out);
int cmd_distance_dev(const cmd_box *box, const double *d_a, const double *d_b, int64_t n,
                     double *d_out);
int cmd_length_all_to_all_dev(const cmd_box *box, const double *d_a, int64_t n,
                              const double *d_b, int64_t m, double *d_out);
int cmd_angle_dev(const cmd_box *box, const double *d_a1, const double *d_a2, const double *d_a3,
                  int64_t n, double *d_out);

/* ---------------------------------------------------------------- jump rates ------------ */
/* Fermi.__call__ / FermiAngle.__call__ (jumprate_generators.py:33-34,42-43) and the legacy
 * kinds.  h_theta may be NULL unless kind == CMD_RATE_FERMI_ANGLE. */
int cmd_rates(int kind, const double h_par[CMD_RATE_NPAR], const double *h_x,
              const double *h_theta, int64_t n, double *h_out);
int cmd_rates_dev(int kind, const double h_par[CMD_RATE_NPAR], const double *d_x,
                  const double *d_theta, int64_t n, double *d_out);

/* ---------------------------------------------------------------- trajectory staging ---- */
/* Page-locked host buffers for trajectory chunks -- what the reference's chunked readers
 * (mdlmc/IO/trajectory_parser.py:296,322: 1000-frame chunks out of HDF5) should fill so that the
 * upload of chunk i+1 overlaps the kernels of chunk i.  Any other host pointer is accepted
 * everywhere too and goes through the library's own staging ring (3 x 8 MiB, page-locked,
 * CMDLMC_B200_STAGE_THREADS host threads, default 4). */
int cmd_host_alloc(size_t bytes, void **out);
int cmd_host_free(void *p);
/* bytes uploaded through the ring / straight from page-locked memory since the library was
 * loaded, and the number of host copy threads (0 before the ring's first use) */
int cmd_staging_stats(uint64_t *staged_bytes, uint64_t *direct_bytes, int *threads);

/* ---------------------------------------------------------------- neighbour topology ---- */
/* A cmd_topo owns, in HBM, the neighbour lists and jump rates of a block of frames:
 *   frame f -> P_f directed pairs stored at [f * stride, f * stride + P_f) of
 *   start i32, dest i32, dist f64, omega f64   (stride = per-frame capacity),
 * in the reference order of NeighborTopology.get_topology_bruteforce (topology.py:55-72):
 * both directions of every pair, row-major, columns ascending; pairs at exactly 0.0 dropped. */

/* mode of cmd_topo_build */
#define CMD_TOPO_BRUTEFORCE 0 /* topology_bruteforce_generator (topology.py:74-78): rebuild every frame */
#define CMD_TOPO_VERLET 1     /* topology_verlet_list_generator (topology.py:80-114) */

/* Creates an empty topology object for n_atoms donor sites; capacity_per_frame = 0 lets the
 * library size it from the first frame. */
int cmd_topo_create(const cmd_box *box, int n_atoms, double cutoff, double buffer, int mode,
                    int rate_kind, const double h_rate_par[CMD_RATE_NPAR],
                    int64_t capacity_per_frame, cmd_topo **out);
void cmd_topo_destroy(cmd_topo *t);
/* Neighbour search path: CMD_PATH_DENSE = one CTA per frame, all N(N-1)/2 pairs out of shared
 * memory (N <= 1024); CMD_PATH_CELL = cell list for large boxes (north_star: "cell-list build and
 * warp-ballot/prefix-sum compaction for large boxes").  CMD_PATH_AUTO (default) picks by size.
 * Both produce identical lists.  Must be called before the first build. */
#define CMD_PATH_AUTO (-1)
#define CMD_PATH_DENSE 0
#define CMD_PATH_CELL 1
int cmd_topo_set_path(cmd_topo *t, int path);
int cmd_topo_path(const cmd_topo *t);
/* Processes the next block of frames (device-resident f64 [nframes][n_atoms][3]); frames are
 * consecutive in time across calls (the Verlet displacement state is carried).  Results of the
 * previous block are overwritten. */
int cmd_topo_build_dev(cmd_topo *t, const double *d_frames, int64_t nframes);
/* Same with host frames: uploads, then builds.  dtype_bytes is 8 (float64) or 4 (float32, as
 * stored by HDF5Trajectory, IO/trajectory_parser.py:324, up-cast on the device).  Page-locked
 * blocks (cmd_host_alloc) are copied from directly and the call returns while the DMA runs;
 * pageable blocks pass through the library's page-locked staging ring (the host copy of one piece
 * overlaps the DMA of the previous one and the kernels of the chunk before). */
int cmd_topo_build(cmd_topo *t, const void *h_frames, int dtype_bytes, int64_t nframes);
/* Streaming form for brute-force blocks: queues the upload and the kernels of the block and returns
 * without waiting.  With two topologies used alternately the upload of block k+1 (and its first
 * chunk in particular, which nothing hides inside one block) overlaps the kernels of block k --
 * the double-buffered chunk upload of a trajectory reader (IO/trajectory_parser.py:296-337 reads
 * the next chunk only after the previous one has been consumed).  The host block must stay
 * untouched until cmd_topo_wait(t) returns; cmd_topo_wait also reports a capacity overflow.
 * cmd_topo_frame_info of such a block waits for that block only, not for later ones.  Needs a
 * topology that has built one block already (the per-frame capacity is known). */
int cmd_topo_build_async(cmd_topo *t, const void *h_frames, int dtype_bytes, int64_t nframes);
int cmd_topo_wait(cmd_topo *t);
/* Donor selection on the device: after this call the host blocks handed to cmd_topo_build /
 * cmd_topo_skip hold ALL n_total atoms of every frame ([nframes][n_total][3], as a trajectory file
 * stores them) and the topology's n_atoms donors are rows h_index[0 .. n_atoms) of each frame --
 * the selection by atom name of trajectory_parser.py:69-71, done by a gather kernel behind the
 * copy instead of a pass over the chunk on the host.  n_total = 0 switches it off again. */
int cmd_topo_set_selection(cmd_topo *t, int n_total, const int *h_index);
/* Frame-block sharding across GPUs: walks a block of frames that precedes this rank's own block
 * through the Verlet displacement / rebuild-decision pass only (topology.py:96-107) and builds
 * just the list of its last rebuild frame -- the state a sequential run has at the block end.
 * No per-frame results are produced.  A no-op (besides counting frames) in brute-force mode. */
int cmd_topo_skip_dev(cmd_topo *t, const double *d_frames, int64_t nframes);
int cmd_topo_skip(cmd_topo *t, const void *h_frames, int dtype_bytes, int64_t nframes);
/* Per-frame results of the last block.  Pointers may be NULL. h_counts: P_f, h_rebuilt: 1 when
 * frame f's list was rebuilt, h_rate_sum: sum of omega over all listed pairs of the frame. */
int cmd_topo_frame_info(const cmd_topo *t, int64_t *h_counts, uint8_t *h_rebuilt,
                        double *h_rate_sum);
int64_t cmd_topo_stride(const cmd_topo *t);
/* After a build returned CMD_ECAPACITY: the number of directed pairs of the largest frame seen,
 * i.e. the per-frame capacity a re-created topology needs at least (0: no overflow so far). */
int64_t cmd_topo_capacity_needed(const cmd_topo *t);
/* Diagnostics of the dense pair kernel's skin list (frame-to-frame reuse of the filter's
 * candidate set, csrc/pairs_dense.cuh): frames that went through it, how many of them rebuilt the
 * list, and the list entries filtered in total.  All zero when the list is not in use. */
int cmd_topo_skin_stats(const cmd_topo *t, int64_t *frames, int64_t *rebuilds, int64_t *list_entries);
/* Number of periodic images, besides the fractionally wrapped vector, that the pair filter of
 * this topology evaluates (general cells; depends on cutoff + buffer against the cell heights). */
int cmd_topo_n_images(const cmd_topo *t);
int64_t cmd_topo_nframes(const cmd_topo *t);
/* The lists of frames [f0, f0 + nf) of the last block in one strided copy per array: host rows of
 * `width` entries each, width >= the largest P_f among them (cmd_topo_frame_info) -- what the
 * reference's per-frame generator protocol (topology.py:80-114) is served from without one
 * round trip per frame.  Pointers may be NULL. */
int cmd_topo_get_block(const cmd_topo *t, int64_t f0, int64_t nf, int64_t width, int *h_start,
                       int *h_dest, double *h_dist, double *h_omega);
/* Copies frame f of the last block to host arrays of at least P_f elements (any may be NULL):
 * the (row, col, data) triple get_topology_bruteforce returns, plus the rates. */
int cmd_topo_get_frame(const cmd_topo *t, int64_t f, int *h_start, int *h_dest, double *h_dist,
                       double *h_omega);
/* Device views of the last block (valid until the next build / destroy). */
int cmd_topo_device_arrays(const cmd_topo *t, const int **d_start, const int **d_dest,
                           const double **d_dist, const double **d_omega,
                           const int **d_counts);
/* Row index of the lists of the last block: int32 [nframes][pitch], pitch = n_atoms + 1 rounded
 * up to a multiple of 4; the pairs that start at site i of frame f are
 * [rowoff[f][i], rowoff[f][i + 1]) of the frame's arrays. */
int cmd_topo_row_offsets(const cmd_topo *t, const int **d_rowoff);
int cmd_topo_n_atoms(const cmd_topo *t);
/* Device pointer of the frames the last block was built from (float64 [nframes][n_atoms][3]). */
int cmd_topo_positions(const cmd_topo *t, const double **d_frames);
/* Tie audit: number of evaluated pairs of the last block whose distance is
 * within 1e-11 relative of cutoff+buffer. */
int64_t cmd_topo_tie_count(const cmd_topo *t);

/* AngleTopology (topology.py:124-167): a second collective variable, the angle
 * extra[group[start]] - donor[start] - donor[dest] of every listed pair (e.g. P-O...O).
 * cmd_topo_set_groups: donor -> index of its extra atom (AngleTopology._determine_groups).
 * cmd_topo_apply_angles: angles of the last block from the extra-atom positions of the same
 * frames, float64/float32 [nframes][n_extra][3]; with a CMD_RATE_FERMI_ANGLE topology the rates
 * of pairs with theta < theta0 become 0 (jumprate_generators.py:42-43). */
int cmd_topo_set_groups(cmd_topo *t, const int *h_group, int n_extra);
int cmd_topo_apply_angles(cmd_topo *t, const void *h_extra_frames, int dtype_bytes);
int cmd_topo_apply_angles_dev(cmd_topo *t, const double *d_extra_frames);
int cmd_topo_get_frame_angles(const cmd_topo *t, int64_t f, double *h_theta);

/* HydroniumTopology (topology.py:170-257), lattice-independent part: per site of every frame of
 * the last block the k listed neighbours with the smallest distance, ascending (k = 4 upstream).
 * Fails when a site has fewer than k listed neighbours, like the reference (topology.py:250). */
int cmd_topo_nearest(cmd_topo *t, int k);
int cmd_topo_near_arrays(const cmd_topo *t, const int **d_start, const int **d_dest,
                         const double **d_dist, const int **d_counts, int64_t *stride);
int cmd_topo_get_frame_nearest(const cmd_topo *t, int64_t f, int *h_dest, double *h_dist);

/* K5 / "jumpstat" (README.md:57-58): histogram of the listed O-O distances of the
 * last block over nbins equal bins of [lo, hi).  Every unordered pair is listed in both
 * directions and counted twice.  Counts are ADDED to the caller's array (int64 on the host,
 * unsigned 64-bit on the device -- e.g. a torch tensor that is all-reduced across GPUs). */
int cmd_topo_distance_histogram(const cmd_topo *t, double lo, double hi, int nbins,
                                int64_t *h_hist);
int cmd_topo_distance_histogram_dev(const cmd_topo *t, double lo, double hi, int nbins,
                                    unsigned long long *d_hist);

/* ---------------------------------------------------------------- KMC ------------------- */
#define CMD_RNG_REPLAY 0 /* host-pregenerated uniforms, bit-exact replay of the reference stream */
#define CMD_RNG_PHILOX 1 /* counter-based Philox4x32-10, key = (seed, replica) */

/* KMCLattice (MDMC.py:28-226) for n_replicas independent replicas on one topology stream.
 * h_lattices: int32 [n_replicas][n_sites], labels 1..P (MDMC.py:68-72). */
int cmd_kmc_create(const cmd_box *box, int n_sites, int n_replicas, const int *h_lattices,
                   double time_step, int rng_mode, uint64_t seed, cmd_kmc **out);
void cmd_kmc_destroy(cmd_kmc *k);
/* Replica sharding across GPUs: local replica r has the GLOBAL id first + r * step, which is what
 * the Philox counter carries -- a replica's random stream does not depend on how many GPUs share
 * the ensemble (rank q of G owns the ids q, q + G, ...: first = q, step = G). */
int cmd_kmc_set_replica_ids(cmd_kmc *k, int first, int step);
/* HydroniumTopology mode (topology.py:170-257): the transitions of a frame are the k nearest listed
 * neighbours of every site (cmd_topo_nearest must have run on the block), and the distance d of a
 * transition is rescaled per replica before the rate is evaluated:
 *   d_relaxed = T(d)   T: CMD_TRANSFORM_RELU (ReLUTransformation, tpar = a, b, d0, left, right) or
 *                         CMD_TRANSFORM_TABLE (InterpolatedTransformation, linear table x -> y)
 *   d' = (1 - r) d + r d_relaxed,  r = min((frame.time - t_last_jump[proton]) / relaxation_time, 1)
 *        (DistanceInterpolator; relaxation_time <= 0: r = 1; a proton that never jumped: r = 1)
 *   omega = rate(d').  Must be called before the first cmd_kmc_advance. */
#define CMD_TRANSFORM_NONE 0
#define CMD_TRANSFORM_RELU 1
#define CMD_TRANSFORM_TABLE 2
int cmd_kmc_set_hydronium(cmd_kmc *k, int rate_kind, const double h_rate_par[CMD_RATE_NPAR],
                          int transform_kind, const double h_tpar[5], const double *h_table_x,
                          const double *h_table_y, int n_table, double relaxation_time,
                          double frame_time_step);
/* float64 [n_replicas][n_sites]: time of the last jump of proton (label - 1); -1 = never. */
int cmd_kmc_get_last_jump_times(const cmd_kmc *k, double *h_tlast);
/* Occupancy histogram (one of the statistics north_star reduces across GPUs): h_counts[s] = number
 * of (replica, consumed frame) pairs that saw site s occupied; *h_replica_frames = sum over
 * replicas of the frames consumed (the normalisation).  Enable before the first advance. */
int cmd_kmc_enable_occupancy(cmd_kmc *k);
int cmd_kmc_get_occupancy(const cmd_kmc *k, int64_t *h_counts, int64_t *h_replica_frames);
/* Replay stream: h_u float64 [n_replicas][n_per_replica]; consumed strictly alternating per
 * event e: h_u[2e] is the TIME SELECTOR -log(1 - r) of the event's np.random.random() draw r
 * (MDMC.py:148), evaluated by the caller with the reference's own log (NumPy) so that no
 * log-implementation difference enters; h_u[2e+1] is the u of its np.random.uniform(0, S) == S*u
 * (MDMC.py:110).  Replaces the previous stream and rewinds every replica's cursor. */
int cmd_kmc_set_replay_stream(cmd_kmc *k, const double *h_u, int64_t n_per_replica);
/* Event log capacity per replica (0 disables logging); (re)starts the log: events are logged
 * from slot 0 again, so a caller drains the log after every cmd_kmc_advance. */
int cmd_kmc_set_event_log(cmd_kmc *k, int64_t max_events_per_replica);
/* Events that did NOT fit the log since the last cmd_kmc_set_event_log, summed over the replicas
 * (the run itself is unaffected; outputs rebuilt from the log would be wrong): 0 is the rule, a
 * caller that reconstructs lattices from the log must check it after every cmd_kmc_advance. */
int64_t cmd_kmc_events_dropped(const cmd_kmc *k);
/* Observables (MDMC.py:179-208, output.py): MSD per axis and covalent autocorrelation every
 * print_frequency frames, reset every reset_frequency frames. d_positions are the donor
 * positions of the frames handed to cmd_kmc_advance. 0/0 disables. */
int cmd_kmc_set_observables(cmd_kmc *k, int reset_frequency, int print_frequency);
/* Frame number 0 of the observables from positions that never pass through the topology: with
 * AngleTopology the reference's _determine_groups (mdlmc/topo/topology.py:142-146) pulls the
 * first trajectory frame through the cached one-shot iterator, so continuous_output
 * (MDMC.py:92-94) yields it as frame 0 at the first event -- the MSD / autocorrelation start
 * there (MDMC.py:193-196) -- while the KMC walks the trajectory from its SECOND frame, which is
 * frame number 1.  h_positions: donor positions [n_sites][3] of that frame.  After
 * cmd_kmc_set_observables, before the first cmd_kmc_advance. */
int cmd_kmc_seed_observables(cmd_kmc *k, const double *h_positions);
/* Consumes the frames of the topology's current block (KMCLattice.continuous_output,
 * MDMC.py:77-99, fastforward_to_next_jump :121-171, move_proton :101-119).  d_positions:
 * f64 [nframes][n_sites][3] of the same block (needed only when observables are enabled). */
int cmd_kmc_advance(cmd_kmc *k, const cmd_topo *t, const double *d_positions);
/* State read-back. Any pointer may be NULL. */
int cmd_kmc_get_state(const cmd_kmc *k, int *h_lattices, double *h_time, int64_t *h_frame,
                      int64_t *h_n_events, int64_t *h_site_updates);
/* Per replica: phase (0 start, 2 running, 3 halted), halt reason (1 replay stream exhausted,
 * 2 no allowed transition -- the reference raises IndexError there) and the number of replay
 * draws consumed from the current stream. */
int cmd_kmc_get_status(const cmd_kmc *k, int *h_phase, int *h_reason, int64_t *h_cursor);
/* Event log of one replica: returns the number of logged events in *n (<= capacity). */
int cmd_kmc_get_events(const cmd_kmc *k, int replica, int64_t capacity, int64_t *n,
                       int64_t *h_frame, double *h_time, int *h_start, int *h_dest,
                       int *h_proton);
/* O-O distance of the jump pair of every logged event of one replica (same order as
 * cmd_kmc_get_events). */
int cmd_kmc_get_event_distances(const cmd_kmc *k, int replica, int64_t capacity, int64_t *n,
                                double *h_dist);
/* Histogram of those distances over all replicas (numerator of the jump probability vs distance
 * statistic); counts are added to the caller's array. */
int cmd_kmc_jump_histogram(const cmd_kmc *k, double lo, double hi, int nbins, int64_t *h_hist);
int cmd_kmc_jump_histogram_dev(const cmd_kmc *k, double lo, double hi, int nbins,
                               unsigned long long *d_hist);
/* Observable rows of one replica: (frame, time, msd_x, msd_y, msd_z, autocorr) per row. */
int cmd_kmc_get_observables(const cmd_kmc *k, int replica, int64_t capacity, int64_t *n,
                            double *h_rows6);
/* rows [first, first + count) of the same table (a caller that drains block by block copies only
 * what is new); cmd_kmc_get_observables with h_rows == NULL returns the number of rows in *n */
int cmd_kmc_get_observable_rows(const cmd_kmc *k, int replica, int64_t first, int64_t count,
                                double *h_rows);
/* Tie audit: decisions (time stepping / selection) within 1e-9 relative of a boundary. */
int64_t cmd_kmc_tie_count(const cmd_kmc *k);
/* Runs with few replicas in exact arithmetic use one CTA per replica and decide move_proton's
 * searchsorted index (MDMC.py:109-111) on a parallel prefix sum; events whose draw lies within
 * rounding distance of an interval end are re-decided with the reference's sequential np.cumsum.
 * Number of such events so far.  (Environment: CMDLMC_B200_KMC_SOLO=0 disables the kernel,
 * CMDLMC_B200_KMC_SELECT_MARGIN=<x> scales the margin; both are read by cmd_kmc_create.) */
int64_t cmd_kmc_selection_fallbacks(const cmd_kmc *k);
/* Tools only: raw 64-bit counter i (0 ties, 1 selection fallbacks, 2.. per-phase cycle counters of
 * builds with -DSOLO_PROFILE, tools/time_kmc_replay.py); -1 on a bad index. */
int64_t cmd_kmc_debug_counter(const cmd_kmc *k, int i);

/* ---------------------------------------------------------------- legacy LMC sweep ------ */
/* PARITY UNPINNED: the legacy engine (LMCHelper.pyx, LMCRoutine.sweep / sweep_with_jumpmatrix) is
 * not in the reference tree; restated from mdlmc/IO/config_parser.py:182-189 and the .bak tests
 * (tests/cython_exts/LMC/test_LMCRoutine.py.bak:11-42, tests/LMC/test_MDMC.py.bak:61-84).
 * One sweep on a frame = P attempts; attempt: pair k uniform in [0, P), u uniform in [0, 1);
 * hop start[k] -> dest[k] iff start occupied, dest empty and u < omega[k] * prob_scale. */
typedef struct cmd_lmc cmd_lmc;
int cmd_lmc_create(int n_sites, int n_replicas, const int *h_lattices, int rng_mode, uint64_t seed,
                   cmd_lmc **out);
void cmd_lmc_destroy(cmd_lmc *k);
/* Replay streams (e.g. GSL gsl_rng_uniform_int(P) / gsl_rng_uniform from a seeded MT19937):
 * h_pick int32 [n_replicas][n], h_acc float64 [n_replicas][n], one entry per attempt. */
int cmd_lmc_set_replay_stream(cmd_lmc *k, const int *h_pick, const double *h_acc, int64_t n);
/* sweep_with_jumpmatrix: counts hops per (start, dest), summed over replicas. */
int cmd_lmc_enable_jump_matrix(cmd_lmc *k, int enable);
/* sweeps_per_frame sweeps on every frame of the topology's current block; prob_scale = dt. */
int cmd_lmc_advance(cmd_lmc *k, const cmd_topo *t, double prob_scale, int sweeps_per_frame);
int cmd_lmc_get_state(const cmd_lmc *k, int *h_lattices, int64_t *h_jumps, int64_t *h_attempts,
                      int64_t *h_sweeps, int *h_halted);
int cmd_lmc_get_jump_matrix(const cmd_lmc *k, int64_t *h_matrix);

/* ---------------------------------------------------------------- multi-GPU ------------- */
/* One process per GPU.  The hot path shards by trajectory-frame block and by independent replica
 * with NO data-path collective; what crosses GPUs is (a) the SUM of the statistics the reference
 * accumulates per run -- MSD sums and autocorrelation of mdlmc/LMC/output.py:17-49, jump counts,
 * the jumpstat histograms -- and (b) the per-frame step lengths dr[N] of
 * mdlmc/topo/topology.py:98 so that every rank can walk the rebuild schedule of the WHOLE
 * trajectory while holding only its own block.  NCCL (NVLink / NVSwitch) is bound with dlopen at
 * the first call; world == 1 turns every collective into an identity. */
/* rank 0: a fresh NCCL unique id (128 bytes) for the host to hand to the other ranks */
int cmd_comm_unique_id(unsigned char h_id[128]);
/* every rank, after cmd_init: builds the communicator on this process's device */
int cmd_comm_init(int rank, int world, const unsigned char h_id[128]);
int cmd_comm_destroy(void);
int cmd_comm_rank(void);
int cmd_comm_world(void);
int cmd_comm_nccl_version(void);
/* in-place sums over all ranks of n_f64 doubles and n_i64 64-bit integers (host buffers) */
int cmd_stats_allreduce(double *h_f64, int64_t n_f64, int64_t *h_i64, int64_t n_i64);
/* the same on device buffers, enqueued on the library stream (no host synchronisation) */
int cmd_stats_allreduce_dev(double *d_f64, int64_t n_f64, int64_t *d_i64, int64_t n_i64);
/* d_recv[r * bytes_per_rank ...] = rank r's d_send, on the library stream */
int cmd_allgather_dev(const void *d_send, void *d_recv, int64_t bytes_per_rank);
/* Frame-block sharding of a Verlet run (topology.py:80-114) WITHOUT walking the coordinates of the
 * frames before the block: every rank computes the step lengths of its own block
 * (cmd_topo_dr_dev), the ranks all-gather them (cmd_allgather_dev), and a fresh topology replays
 * the rebuild schedule of the preceding frames from those numbers alone (cmd_topo_skip_dr_dev),
 * then takes the coordinates of just two frames -- the last rebuild frame and the frame before
 * the block (cmd_topo_seed_dev).  The state equals that of a sequential run at the block start. */
int cmd_topo_dr_dev(const cmd_topo *t, const double *d_frames, int64_t nframes,
                    const double *d_prev, double *d_dr);
int cmd_topo_skip_dr_dev(cmd_topo *t, const double *d_dr, int64_t nframes, int64_t *h_last_rebuild);
int cmd_topo_seed_dev(cmd_topo *t, const double *d_frame_rebuild, const double *d_frame_prev);
/* ADDS (number of listed directed pairs, sum of the listed rates) of the current block to
 * d_out[0..1] on the device: the block statistics a frame-block shard contributes per step */
int cmd_topo_block_stats_dev(const cmd_topo *t, double *d_out);

#ifdef __cplusplus
}
#endif
#endif
