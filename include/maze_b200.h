/*
 * maze_b200.h -- C-ABI of the B200 LOKI re-segmentation stage.
 *
 * The reference (MAZE-IPP) is pure Python and has no FFI; its hot path is a chain of Python
 * callables invoked once per vignette (paths relative to the reference root):
 *
 *   threshold            image > threshold_brighter                 maze_ipp/loki/pipeline.py:649
 *   bool cast            np.asarray(pred, dtype=bool)               maze_ipp/loki/pipeline.py:405
 *   isotropic_erosion    edt(image) > radius                        maze_ipp/isotropic.py:8-36
 *   isotropic_dilation   edt(image == 0) < radius                   maze_ipp/isotropic.py:39-67
 *   isotropic_opening    erosion, dilation                          maze_ipp/isotropic.py:70-98
 *   isotropic_closing    dilation, erosion                          maze_ipp/isotropic.py:101-129
 *   label                skimage.measure.label(bool), 8-conn        maze_ipp/loki/pipeline.py:430-433
 *   clear_border         clear_border(labels, out=labels)           maze_ipp/loki/pipeline.py:435-439
 *   remove_small_objects remove_small_objects(labels, min_size, out=labels)   :442-448
 *   merge_labels         merge_labels(labels, max_distance, labels_out=labels)
 *                                                                   maze_ipp/merge_labels.py:29-113
 *   regionprops          FindRegions / ImageProperties + RegionProperties reads
 *                                                                   maze_ipp/loki/pipeline.py:589-625, 653-654
 *
 * Every entry point below replaces one of those callables for a whole PACKED BATCH of
 * vignettes.  All pointers are DEVICE pointers unless the name ends in _host; the caller owns
 * every buffer (the library never allocates); `stream` is a cudaStream_t passed as void*.
 * Calls are asynchronous on `stream` and re-entrant across streams.  Return value: MAZE_OK or
 * a negative MAZE_ERR_* code (maze_error_string() gives the text of the last CUDA error seen
 * by the calling thread).
 *
 * Batch geometry
 * --------------
 * A batch is n_img vignettes.  Vignette i is h x w pixels; its pixels live row-major and
 * contiguous (numpy C order, no row padding) at element offset pix_off in every per-pixel
 * array of the batch (uint8 image, uint8/bool mask, int32 labels, int32 scratch).  pix_off must
 * be a multiple of 16.  Binary images are additionally kept as BIT PLANES: row y of vignette i
 * is wpr = ceil(w/32) uint32 words at word offset word_off + y*wpr, pixel x is bit (x & 31) of
 * word (x >> 5); bits at x >= w are always stored as 0.  The batch is cut into TILES of
 * MAZE_TILE_WORDS consecutive words of one vignette (a tile never straddles vignettes); one
 * CTA processes one tile.  tile0 is the index of the vignette's first tile.
 */
#ifndef MAZE_B200_H
#define MAZE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAZE_OK 0
#define MAZE_ERR_CUDA (-1)      /* a CUDA call or launch failed */
#define MAZE_ERR_TYPEERROR (-2) /* merge_labels: the reference raises TypeError here (merge_labels.py:19-20) */
#define MAZE_ERR_BADARG (-3)
#define MAZE_ERR_CAPACITY (-4)  /* a caller-provided table is too small */

#define MAZE_TILE_WORDS 256
#define MAZE_MAX_DISK_RADIUS 32 /* bit-plane morphology handles d2 thresholds < 33*33; beyond: maze_edt_sq */

typedef struct maze_vignette {
    int64_t pix_off;  /* element offset of pixel (0,0) in the per-pixel arrays */
    int64_t word_off; /* word offset of row 0 in a bit plane */
    int32_t h, w;
    int32_t wpr;      /* ceil(w / 32) */
    int32_t tile0;    /* index of the first tile of this vignette */
} maze_vignette_t;    /* 32 bytes */

typedef struct maze_tile {
    int32_t img;   /* vignette index */
    int32_t word0; /* first word (relative to the vignette's word_off) covered by the tile */
} maze_tile_t;

/* Feature table: one row of MAZE_NFEAT doubles per object.  Same column layout as the oracle. */
#define MAZE_NFEAT 64
#define MAZE_F_LABEL 0
#define MAZE_F_AREA 1
#define MAZE_F_BBOX 2         /* 4: min_row, min_col, max_row (excl), max_col (excl) */
#define MAZE_F_CENTROID 6     /* 2: row, col */
#define MAZE_F_MU 8           /* 16: central moments mu[p*4+q], p = row power, q = col power */
#define MAZE_F_NU 24          /* 16: normalised central moments */
#define MAZE_F_HU 40          /* 7 */
#define MAZE_F_EIG 47         /* 2: inertia tensor eigenvalues, descending */
#define MAZE_F_AXIS_MAJOR 49
#define MAZE_F_AXIS_MINOR 50
#define MAZE_F_ECC 51
#define MAZE_F_ORIENT 52
#define MAZE_F_IMIN 53
#define MAZE_F_IMAX 54
#define MAZE_F_IMEAN 55
#define MAZE_F_FRAC_INVALID 56 /* mean(intensity == 0) inside the region, loki/pipeline.py:617 */
#define MAZE_F_IMAGE 57        /* vignette index of the object */
#define MAZE_F_T00 58          /* inertia tensor */
#define MAZE_F_T01 59
#define MAZE_F_T11 60

/* Shape table (maze_label_shape): one row of MAZE_NSHAPE doubles per object, same row order as the feature table. */
#define MAZE_NSHAPE 8
#define MAZE_S_PERIMETER 0   /* skimage.measure.perimeter(region.image, neighborhood=4) */
#define MAZE_S_FILLED_AREA 1 /* region.image with holes filled (ndi.binary_fill_holes, full 3x3 structure) */
#define MAZE_S_EULER 2       /* skimage.measure.euler_number(region.image, connectivity=2) */
#define MAZE_S_N1 3          /* border pixels of weight 1, sqrt(2), (1 + sqrt(2)) / 2: the exact integers behind */
#define MAZE_S_N2 4          /*   the perimeter */
#define MAZE_S_N3 5
#define MAZE_S_CONVEX_AREA 6 /* np.sum(skimage.morphology.convex_hull_image(region.image)) */

/* Per-object integer accumulators (regionprops pass 1): MAZE_NACC uint64 per object followed,
 * in a second array, by MAZE_NEXT int32 extrema per object. */
#define MAZE_NACC 12
#define MAZE_NEXT 8

#define MAZE_RP_HIGH_ORDER 1 /* also fill mu/nu[p,q] with p+q > 3 (second pass over the labels) */
#define MAZE_RP_RUNS 2       /* labels AND bits given, labels are constant along the runs of bits (output of
                                maze_label, also after the label filters): run-based reduction */

const char *maze_error_string(void);
int maze_version(void);

/* loki/pipeline.py:649 (and :405 with t_int = 0): bit = (pixel > t_int).  The caller folds the
 * float threshold to the integer t_int = floor(threshold) clipped to [-1, 255].
 * flags[i] receives bit0 = "mask has a 1", bit1 = "mask has a 0" (zeroed by this call). */
int maze_threshold_pack(const uint8_t *image, const maze_vignette_t *vig, int n_img,
                        const maze_tile_t *tiles, int n_tiles, int t_int,
                        uint32_t *bits, uint32_t *flags, void *stream);

/* One thresholded-EDT pass on bit planes.  invert = 0: isotropic_erosion (isotropic.py:35-36),
 * out = [d2(in) > t]; invert = 1: isotropic_dilation (isotropic.py:66-67), out = [d2(in == 0) <= t].
 * t is the integer squared-distance threshold folded from the float radius by the caller
 * (erosion: max{k : sqrt(k) <= radius}; dilation: max{k : sqrt(k) < radius}; -1 = none) and must
 * be < (MAZE_MAX_DISK_RADIUS+1)^2.  flags_in are the flags of `in`; flags_out (zeroed by this
 * call) receive the flags of `out`.  in != out. */
int maze_morph_pass(const uint32_t *in, uint32_t *out, const maze_vignette_t *vig, int n_img,
                    const maze_tile_t *tiles, int n_tiles, int t, int invert,
                    const uint32_t *flags_in, uint32_t *flags_out, void *stream);

/* Plain binary morphology with a footprint (skimage.morphology.binary_erosion / binary_dilation as called through
 * binary_opening / binary_closing at loki/pipeline.py:408-427): the footprint is symmetric and row-convex, given as
 * half chords w[|dy|] for |dy| <= R (row dy holds the pixels |dx| <= w[|dy|]); w must not increase with |dy|.
 * disk(r) and disk(r, decomposition="crosses") both collapse to one such footprint (erosion by a sequence of
 * elements is the erosion by their Minkowski sum).  maze_footprint_register returns an id >= 0 (the same id for the
 * same table; negative = error); pass MAZE_FOOTPRINT_T(id) wherever a pass threshold `t` is expected
 * (maze_morph_pass, the pass tables of maze_front_chain / maze_vignette_stage / maze_stage_step): invert = 0 is then
 * the erosion (pixels outside the image count as foreground, skimage's border_value=True), invert = 1 the dilation
 * (outside = background); scipy's EDT phantom pixel does not apply. */
#define MAZE_FOOTPRINT_T(id) (-2 - (id))
int maze_footprint_register(int R, const int32_t *w);

/* The same pass for LARGE radii (squared-distance thresholds up to 254^2, or a registered footprint): a separable
 * pair of kernels -- vertical distance to the nearest 0 per pixel (one byte, capped at R + 1), then a row test
 * against the disk's half heights -- whose cost grows with R instead of R^2; results are identical to
 * maze_morph_pass.  cta_off_a[i] (n_img + 1 int64, device) = prefix sum of ceil(h / 32) * ceil(wpr / 8) over the
 * vignettes, n_cta_a its total; cta_off_b / n_cta_b the same for ceil(h / 8) * ceil(w / 256).  g_scratch: one byte per pixel; plane_scratch: one bit plane (!= in, out). */
int maze_morph_pass_wide(const uint32_t *in, uint32_t *out, const maze_vignette_t *vig, int n_img,
                         const int64_t *cta_off_a, long long n_cta_a, const int64_t *cta_off_b, long long n_cta_b,
                         int t, int invert, const uint32_t *flags_in, uint32_t *flags_out, uint8_t *g_scratch,
                         uint32_t *plane_scratch, void *stream);

/* bit plane -> one byte per pixel (numpy bool). */
int maze_unpack_mask(const uint32_t *bits, const maze_vignette_t *vig, int n_img,
                     const maze_tile_t *tiles, int n_tiles, uint8_t *mask, void *stream);

/* Exact squared Euclidean distance transform (scipy.ndimage.distance_transform_edt as used at
 * isotropic.py:35,66 and merge_labels.py:17,22): d2[p] = squared distance from pixel p to the
 * nearest 0 bit of `bits`, 0 where the bit is 0; a plane without any 0 bit behaves as if a
 * single 0 sat at (-1, 0).  invert = 1 transforms the complement plane (distance to the nearest 1
 * bit).  d2 is per-pixel int32 and doubles as the column-pass scratch; max_h / max_w are the
 * largest vignette height / width of the batch; flag_scratch holds n_img uint32. */
int maze_edt_sq(const uint32_t *bits, const maze_vignette_t *vig, int n_img, int max_h, int max_w,
                int invert, int32_t *d2, uint32_t *flag_scratch, void *stream);

/* out bit = (d2 > t) (greater = 1) or (d2 <= t) (greater = 0): the compare of isotropic.py:36,67
 * for radii beyond MAZE_MAX_DISK_RADIUS. */
int maze_compare_pack(const int32_t *d2, const maze_vignette_t *vig, int n_img,
                      const maze_tile_t *tiles, int n_tiles, int t, int greater,
                      uint32_t *bits, uint32_t *flags, void *stream);

/* loki/pipeline.py:430-433: 8-connected components, labels 1..N in raster order of each
 * component's first pixel, background 0, int32.  parent: per-pixel int32 scratch (touched only
 * at run starts).  tile_scan: n_tiles + 1 int32 scratch.  lab_off: n_img + 1 int32, receives the
 * exclusive prefix sum of the per-vignette label counts (object index of (i, l) = lab_off[i] + l - 1). */
int maze_label(const uint32_t *bits, const maze_vignette_t *vig, int n_img,
               const maze_tile_t *tiles, int n_tiles, int32_t *parent, int32_t *labels,
               int32_t *tile_scan, int32_t *lab_off, void *stream);

/* loki/pipeline.py:435-439 and :442-448, in place, no renumbering.  Labels must lie in
 * [0, lab_off[i+1]-lab_off[i]]; obj_scratch holds lab_off[n_img] int32 (n_obj_cap entries are
 * cleared by the call). */
int maze_clear_border(int32_t *labels, const maze_vignette_t *vig, int n_img,
                      const maze_tile_t *tiles, int n_tiles, const int32_t *lab_off,
                      int32_t *obj_scratch, int n_obj_cap, void *stream);
int maze_remove_small_objects(int32_t *labels, const maze_vignette_t *vig, int n_img,
                              const maze_tile_t *tiles, int n_tiles, const int32_t *lab_off,
                              int32_t *obj_scratch, int n_obj_cap, int64_t min_size, void *stream);

/* largest label of every vignette (for label images that did not come from maze_label). */
int maze_max_label(const int32_t *labels, const maze_vignette_t *vig, int n_img,
                   const maze_tile_t *tiles, int n_tiles, int32_t *max_label, void *stream);

/* Per-label regionprops (loki/pipeline.py:589-625, 653-654).  labels may be NULL, in which case
 * `bits` is read as a label image with the single label 1 (ImageProperties semantics).  image may
 * be NULL (no intensity columns).  acc: n_obj_cap * MAZE_NACC uint64; ext: n_obj_cap * MAZE_NEXT
 * int32; table: n_obj_cap * MAZE_NFEAT doubles.  Rows of absent labels get area 0 and NaN.
 * acc_base (optional, n_img int32): rows of vignettes with acc_base[i] >= 0 are not touched (they
 * come from maze_props_finish_staged). */
int maze_regionprops(const int32_t *labels, const uint32_t *bits, const uint8_t *image,
                     const maze_vignette_t *vig, int n_img, const maze_tile_t *tiles, int n_tiles,
                     const int32_t *lab_off, int n_obj_cap, unsigned long long *acc, int32_t *ext,
                     double *table, int flags, const int32_t *acc_base, void *stream);

/* Shape features read by CalculateZooProcessFeatures(region, meta, prefix="object_") (loki/pipeline.py:625, 654)
 * through skimage's RegionProperties: perimeter, euler_number, filled_area, convex_area (SURVEY.md section 8, rows
 * a10 / f1).
 * One row of MAZE_NSHAPE doubles per row of `table` (the finished feature table of maze_regionprops /
 * maze_props_finish_staged, which supplies label, vignette and bounding box); rows with area 0 get NaN.
 * labels may be NULL: `bits` is then the single region (ImageProperties semantics, loki/pipeline.py:653).
 * runs = 1: labels AND bits given and the labels are constant along the runs of bits (as for MAZE_RP_RUNS): the
 * object planes are then cut from the bit plane, a fraction of the loads.
 * pool: n_slabs slabs of slab_words uint32 scratch, slab_words >= 2 * (max_h + 2) * ceil((max_w + 2) / 32) for the
 * tallest / widest bounding box of the batch (used by the objects whose planes exceed shared memory, one CTA per
 * slab; required).  max_h: tallest vignette of the batch (sizes the hull storage of those objects; objects taller
 * than 12 499 rows get NaN for the convex area).  Coordinates are packed in 16 bits: sides up to 32 767 pixels.  work_counter: two int32
 * (cleared by the call). */
int maze_label_shape(const int32_t *labels, const uint32_t *bits, const maze_vignette_t *vig,
                     const double *table, int n_obj, uint32_t *pool, long long slab_words, int n_slabs,
                     int max_h, int runs, int32_t *work_counter, double *shape, void *stream);

/* maze_ipp/merge_labels.py:29-113 for every vignette of the batch, one CTA per vignette.
 * labels: read by the loop; labels_out: written (pass the same pointer for the pipeline's aliased
 * call, loki/pipeline.py:452-457).  index/index_off: optional explicit label lists (index_off has
 * n_img + 1 entries) or NULL for "sorted positive labels".  lab_off bounds the label values as in
 * maze_clear_border.  have_max = 0 restates max_distance=None.  d2a, d2b, d2c: per-pixel int32 scratch;
 * obj_scratch: 7 * n_obj_cap int32 (label list and per-label minima in the first 2 * n_obj_cap; label boxes and flags
 * of the windowed kernel behind).  merge_dist (n_obj_cap doubles, optional) / n_merge (n_img):
 * the distances at which labels were merged.  index_state (2 * n_img int32): the length of the
 * label list the loop started from and how many entries it popped (the list itself, popped entries
 * first, is left in obj_scratch + 2 * lab_off[i]).  status (n_img int32): MAZE_OK or
 * MAZE_ERR_TYPEERROR per vignette.  order (optional, n_img int32): CTA b works on vignette order[b] (put the
 * largest vignettes first: one CTA per vignette, the longest ones decide the makespan). */
int maze_merge_labels(const int32_t *labels, int32_t *labels_out, const maze_vignette_t *vig, int n_img,
                      const int32_t *lab_off, int n_obj_cap, const int32_t *index, const int32_t *index_off,
                      int have_max, double max_distance, double path_tolerance,
                      int32_t *d2a, int32_t *d2b, int32_t *d2c, int32_t *obj_scratch,
                      double *merge_dist, int32_t *n_merge, int32_t *index_state, int32_t *status,
                      const int32_t *order, void *stream);

/* The same for a SUBSET of the vignettes and with a choice of how many CTAs work on one vignette: the call handles
 * the n_order vignettes order[0 .. n_order) (order = NULL: all n_img) with cluster_size CTAs each -- 1, or 8: a
 * thread-block cluster (cluster barrier between the steps, reductions through distributed shared memory) that gives a
 * large vignette eight times the memory parallelism of one CTA.  Vignettes that are not listed are not touched
 * (n_merge / status / index_state keep what the caller put there).
 *
 * index = NULL with have_max = 1 -- the pipeline's call -- runs the WINDOWED kernel (maze_merge_win.cu; cluster_size is
 * ignored, MAZE_MERGE_WINDOWED=0 in the environment selects the whole-image kernel): a preparation kernel finds the
 * bounding box of every label with the whole GPU, then one 1024-thread CTA per vignette runs the loop and every
 * iteration only touches the distance window of the label it pops; the loop ends without a distance map when the
 * popped label's minimum of distmap already exceeds max_distance^2.  Same results, bit for bit. */
int maze_merge_labels_ex(const int32_t *labels, int32_t *labels_out, const maze_vignette_t *vig, int n_img,
                         const int32_t *lab_off, int n_obj_cap, const int32_t *index, const int32_t *index_off,
                         int have_max, double max_distance, double path_tolerance,
                         int32_t *d2a, int32_t *d2b, int32_t *d2c, int32_t *obj_scratch,
                         double *merge_dist, int32_t *n_merge, int32_t *index_state, int32_t *status,
                         const int32_t *order, int n_order, int cluster_size, void *stream);

/* Synthetic LOKI-shaped vignettes for the benchmark: dark noisy background plus
 * 1-6 anisotropic Gaussian blobs per vignette, counter-based RNG keyed by (seed, vignette, pixel). */
int maze_synth_vignettes(uint8_t *image, const maze_vignette_t *vig, int n_img,
                         const maze_tile_t *tiles, int n_tiles, uint64_t seed, int64_t img_index0,
                         void *stream);

/* Vignette-resident fused stage: threshold -> n_pass (<= 4) thresholded-EDT passes -> label() ->
 * per-label regionprops accumulators, one CTA per vignette with the bit planes, the union-find and
 * the accumulators in shared memory.  Writes the final bit plane, the mask bytes, the int32 label
 * image, n_labels[i] and -- into the staging arrays -- one accumulator row per label.
 * image: bytes that are thresholded (pixel > t_int); intensity: bytes the intensity features are
 * taken from (may be the same pointer, or NULL).
 * img_list (device) holds the vignette indices grouped into MAZE_FUSED_CLASSES size classes (h*wpr
 * words <= MAZE_FUSED_CAPS[c]); class c is img_list[class_off_host[c] .. class_off_host[c+1]).
 * Larger vignettes must go through the per-operator entry points above.  pass_t_host /
 * pass_invert_host: the d2 threshold and erosion(0)/dilation(1) flag of every pass, as for
 * maze_morph_pass.  flags: MAZE_RP_HIGH_ORDER, MAZE_FUSED_NO_PROPS.
 * fallback[i] = 1 marks a listed vignette with more word runs than union-find slots (nothing was
 * written for it; use the per-operator path).  acc_base[i] receives the first staging row of
 * vignette i (label l is row acc_base[i] + l - 1) or -1 when it was not staged (fallback, or the
 * stage_cap rows of acc_stage [MAZE_NACC u64 each] / hi_stage [8 doubles] / ext_stage [MAZE_NEXT
 * int32] were used up: use maze_regionprops for those).  stage_counter: one int32 of scratch. */
#define MAZE_FUSED_NO_PROPS 4 /* flag: labels only, no accumulators are staged (acc_base[i] = -1) */
#define MAZE_FUSED_CLASSES 8
#define MAZE_FUSED_CAPS {1024, 2560, 4096, 6144, 9216, 13312, 19456, 28320}
int maze_vignette_stage(const uint8_t *image, const uint8_t *intensity, const maze_vignette_t *vig,
                        const int32_t *img_list, const int32_t *class_off_host, int t_int, int n_pass,
                        const int32_t *pass_t_host, const int32_t *pass_invert_host, int flags,
                        uint32_t *bits, uint8_t *mask, int32_t *labels, int32_t *n_labels, int32_t *fallback,
                        int32_t *acc_base, int32_t *stage_counter, int stage_cap,
                        unsigned long long *acc_stage, double *hi_stage, int32_t *ext_stage, void *stream);

/* ---- Band pipeline (maze_bands.cu): the same chain as maze_vignette_stage, cut into uniform pieces ----------
 * A BAND is a group of consecutive rows [y0, y1) of one vignette; together with the halo rows it recomputes
 * (halo = sum of the pass radii on each side, none when the band is the whole vignette) it holds at most
 * MAZE_BAND_PLANE_WORDS words.  All bands of a vignette have rpb rows (the last one may be shorter) and are
 * consecutive in the band list; band_off[i] .. band_off[i+1] are the bands of vignette i (none: the vignette is
 * not handled by this entry point -- a row plus halo does not fit, or a side >= 65536).
 * Results: the final bit plane; the RUN LIST -- every maximal horizontal run of foreground pixels as
 * {y, x0, x1 (inclusive), label}, the runs of a band contiguous and in raster order at runs[band_out[b].base ..
 * + n_runs) -- which is also the compact form of the label image that crosses PCIe; n_labels / fallback /
 * acc_base / staged accumulators exactly as for maze_vignette_stage; and, when mask and labels are given, the
 * dense bool mask and int32 label image (both arrays, total_px elements each -- a multiple of 16 --, are zero
 * filled on a forked stream next to the band kernel, the runs are stored on top; needs run_pix, one uint32 per
 * run slot, and total_px < 2^32).
 * counters: 8 int32 (cleared by the call): [0] staging rows used, [1] runs used (may exceed run_cap: the bands
 * that did not fit are flagged), [2..4] vignettes that needed the larger run tables.  big_list: 3 * n_img
 * int32 scratch.
 * FRAMES: vignettes with at least huge_px pixels (huge_host: n_huge pairs {vignette index, number of its bands}) are
 * labelled by a sequence of global-memory kernels (union-find on run ids with atomics) instead of the per-vignette
 * CTA; gl_scratch: 2 * run_cap + n_bands + 16 int32.  n_huge = 0 or gl_scratch = NULL switches this off.
 * band_done (n_img int32 scratch, cleared by the call; may be NULL): with it, the LAST band CTA of a vignette to finish
 * labels the vignette in place inside the band kernel (run tables up to 4096 runs / 2048 rows / 64 bands, the rest goes to
 * the list kernel); without it a labelling kernel of its own runs behind the band kernel.
 * clear_border / min_area: the label filters of loki/pipeline.py:435-448 applied ON THE RUN LIST by the labelling kernel
 * (labels touching the outermost rows / columns, labels with fewer than min_area pixels: their runs get label 0, their
 * rows are emptied; no renumbering); not available together with frames (n_huge > 0), nor is MAZE_BAND_SINGLE_REGION.
 * flags: MAZE_RP_HIGH_ORDER, MAZE_FUSED_NO_PROPS, MAZE_BAND_SINGLE_REGION.
 * fallback[i] = 1: nothing valid was produced for vignette i (more runs than slots, run buffer full, or scipy's
 * phantom pixel applies to a multi-band vignette): use the per-operator entry points for it. */
#define MAZE_BAND_PLANE_WORDS 6144
#define MAZE_BAND_SINGLE_REGION 8 /* flag: no labelling, the whole mask of a vignette is label 1 (ImageProperties,
                                     loki/pipeline.py:653); n_labels is 1 for every vignette, also for an empty mask */
typedef struct maze_band { int32_t img, y0, y1, rpb; } maze_band_t;
typedef struct maze_band_out { int32_t base, n_runs, zflags, reserved; } maze_band_out_t;
typedef struct maze_run { uint16_t y, x0, x1, label; } maze_run_t;
typedef struct maze_run_stat { uint32_t isum; uint16_t zeros; uint8_t vmin, vmax; } maze_run_stat_t;
int maze_band_stage(const uint8_t *image, const uint8_t *intensity, const maze_vignette_t *vig, int n_img,
                    const maze_band_t *bands, int n_bands, const int32_t *band_off, int t_int, int n_pass,
                    const int32_t *pass_t_host, const int32_t *pass_invert_host, int halo, int flags,
                    uint32_t *bits, maze_run_t *runs, uint32_t *run_pix, maze_run_stat_t *run_stats, int run_cap,
                    maze_band_out_t *band_out, uint8_t *mask, int32_t *labels, int32_t *n_labels,
                    int32_t *fallback, int32_t *acc_base, int32_t *counters, int32_t *big_list, int stage_cap,
                    unsigned long long *acc_stage, double *hi_stage, int32_t *ext_stage, long long total_px,
                    const int32_t *huge_host, int n_huge, long long huge_px, int32_t *gl_scratch, int32_t *band_done,
                    int clear_border, long long min_area, void *stream);

/* Feature rows from the staged accumulators: row lab_off[i] + l - 1 of table for every vignette with
 * acc_base[i] >= 0 (the others are left to maze_regionprops). */
int maze_props_finish_staged(const unsigned long long *acc_stage, const double *hi_stage, const int32_t *ext_stage,
                             const int32_t *acc_base, const int32_t *lab_off, int n_img, int n_obj,
                             int has_intensity, int flags, double *table, void *stream);

/* lab_off[0..n_img] = exclusive prefix sum of n_labels[0..n_img). */
int maze_count_scan(const int32_t *n_labels, int n_img, int32_t *lab_off, void *stream);

/* threshold -> n_pass thresholded-EDT passes (d2 thresholds < 33^2) -> label -> mask bytes in one call, with
 * the per-operator kernels above (for vignettes too large for maze_vignette_stage).  plane_a / plane_b:
 * bit-plane scratch, flags_a / flags_b: n_img uint32 each; *final_plane_host receives the plane (a or b)
 * that holds the final mask.  Other arguments as for maze_label / maze_unpack_mask. */
int maze_front_chain(const uint8_t *image, const maze_vignette_t *vig, int n_img, const maze_tile_t *tiles,
                     int n_tiles, int t_int, int n_pass, const int32_t *pass_t_host, const int32_t *pass_invert_host,
                     uint32_t *plane_a, uint32_t *plane_b, uint32_t *flags_a, uint32_t *flags_b, int32_t *parent,
                     int32_t *labels, int32_t *tile_scan, int32_t *lab_off, uint8_t *mask,
                     uint32_t **final_plane_host, void *stream);

/* One asynchronous step of the fused stage in a single call (what stage.LokiSegmentationStage issues per
 * batch): counters zeroed, maze_vignette_stage on `lane_stream`, maze_front_chain for the left_n vignettes
 * that are too large for it on `side_stream` (forked / joined with events), maze_count_scan,
 * maze_props_finish_staged, maze_regionprops for the oversize vignettes (trailing on the side stream), and
 * the copy of counts[3*n_img] (n_labels | fallback | acc_base), of the object total and (band pipeline) of the
 * number of runs used into pinned counts_host[3*n_img + 2].  left_vig / left_tiles describe the oversize vignettes as their own batch (same
 * offsets), left_idx (device) holds their indices in the full batch, left_tiles_full their tiles with
 * full-batch vignette indices.  pass_t / pass_invert: as for maze_vignette_stage (no t == 0 passes).
 * scratch_*: plane (total words), flags (2*left_n), parent (per pixel), tile_scan (left_n_tiles+1),
 * lab_off (left_n+1), acc / ext (stage_cap rows). */
typedef struct maze_step_args {
    const maze_vignette_t *vig;
    const int32_t *img_list;
    const maze_vignette_t *left_vig;
    const maze_tile_t *left_tiles;
    const int32_t *left_idx;
    const maze_tile_t *left_tiles_full;
    const uint8_t *image, *intensity;
    uint32_t *bits;
    uint8_t *mask;
    int32_t *labels, *counts, *lab_off, *stage_counter;
    unsigned long long *acc_stage;
    double *hi_stage;
    int32_t *ext_stage;
    double *table;
    uint32_t *scratch_plane, *scratch_flags;
    int32_t *scratch_parent, *scratch_tile_scan, *scratch_lab_off;
    unsigned long long *scratch_acc;
    int32_t *scratch_ext;
    int32_t *counts_host;
    /* band pipeline (bands != NULL: maze_band_stage runs in place of maze_vignette_stage) */
    const maze_band_t *bands;
    const int32_t *band_off;
    maze_run_t *runs;
    uint32_t *run_pix; /* run_cap uint32: element offset of every run's first pixel (dense outputs only) */
    maze_run_stat_t *run_stats;
    maze_band_out_t *band_out;
    int32_t *band_counters; /* 8 int32 */
    int32_t *big_list;      /* 3 * n_img int32 */
    const int32_t *huge_host; /* HOST: n_huge pairs {vignette, bands} of the frames (global-memory labelling) */
    int32_t *gl_scratch;      /* 2 * run_cap + n_bands + 16 int32 (or NULL) */
    int32_t *band_done;       /* n_img int32 (or NULL) */
    int32_t class_off[MAZE_FUSED_CLASSES + 1];
    int32_t pass_t[4], pass_invert[4];
    int32_t n_img, left_n, left_n_tiles, left_n_tiles_full, t_int, n_pass, flags, stage_cap;
    int32_t n_bands, halo, run_cap, step_flags; /* step_flags: MAZE_STEP_COMPACT */
    int64_t total_px;                           /* elements of mask / labels (band pipeline, dense outputs) */
    int64_t huge_px;                            /* vignettes with >= huge_px pixels are frames */
    int64_t min_area;                           /* label filters on the run list (band pipeline) */
    int32_t n_huge, clear_border;
} maze_step_args_t;
#define MAZE_STEP_COMPACT 1 /* band pipeline: no dense mask / label image for the band vignettes (run list only) */
int maze_stage_step(const maze_step_args_t *args_host, void *lane_stream, void *side_stream);

/* The same step as a CUDA graph (a batch that is one frame is some twenty small launches: issuing them takes the host
 * longer than the GPU needs for the frame).  *exec == NULL: the step is captured on the lane stream, instantiated and
 * launched; the handle and the number of kernel launches inside are returned.  Otherwise the graph is launched.  A graph
 * replays the arguments it was captured with -- key the handles by the argument block (and the band plan behind
 * huge_host).  Steps with leftover vignettes (left_n > 0) or under maze_prof_enable run plainly: *n_launches = -1. */
int maze_stage_step_graph(const maze_step_args_t *a, void *lane_stream, void *side_stream, void **exec, int *n_launches);
int maze_graph_destroy(void *exec);

/* HOST helper: copies n host arrays (srcs[i], nbytes[i] bytes) to dst + dst_off[i] with n_threads threads.
 * Used to pack the vignettes of a batch into one pinned staging buffer (one upload per batch). */
int maze_host_pack(const void *const *srcs_host, const int64_t *nbytes_host, const int64_t *dst_off_host, int n,
                   void *dst_host, int n_threads);

/* Asynchronous form of maze_host_pack: the copy threads start at once, the call returns a job handle (NULL: bad
 * argument) and maze_host_pack_wait joins them (and frees the job).  The descriptor arrays are copied into the job;
 * the source arrays and dst_host must stay alive until the wait returns. */
void *maze_host_pack_start(const void *const *srcs_host, const int64_t *nbytes_host, const int64_t *dst_off_host,
                           int n, void *dst_host, int n_threads);
int maze_host_pack_wait(void *job);

/* HOST helpers of the compact result transport.  In compact mode the label image of a vignette crosses PCIe as the
 * run list of maze_band_stage (8 bytes per run) instead of 5 bytes per pixel; these functions expand it on the
 * host into what the reference's stage returns (bool mask + int32 labels, loki/pipeline.py:459) or into the
 * crop of one object (what FindRegions / ExtractROI consume, loki/pipeline.py:589-602).  runs / band_out are
 * HOST copies of the device arrays.  maze_host_expand: n vignettes, vignette k = bands band_lo[k] .. band_hi[k],
 * h[k] x w[k] pixels, written to mask_dst[k] / label_dst[k] (either array of pointers may be NULL), n_threads
 * threads.  maze_host_expand_crop: rows [r0, r1) x columns [c0, c1) of one vignette (inside the image; rpb = rows
 * per band of that vignette); only_label > 0 keeps that object alone.  MAZE_ERR_CAPACITY: the vignette has no run
 * list (it was flagged as fallback). */
int maze_host_expand(const maze_run_t *runs_host, const maze_band_out_t *band_out_host, const int32_t *band_lo_host,
                     const int32_t *band_hi_host, const int32_t *h_host, const int32_t *w_host, int n,
                     uint8_t *const *mask_dst_host, int32_t *const *label_dst_host, int n_threads);
int maze_host_expand_crop(const maze_run_t *runs_host, const maze_band_out_t *band_out_host, int band_lo, int band_hi,
                          int rpb, int r0, int r1, int c0, int c1, int only_label, uint8_t *mask_dst_host,
                          int32_t *label_dst_host);

/* Launch accounting and optional per-kernel timing (CUDA events on the launching stream).
 * maze_launch_count: kernels launched by this library since load.  With maze_prof_enable(1) every
 * launch is bracketed by an event pair; maze_prof_collect waits for them and ADDS elapsed
 * milliseconds / launch counts per kernel id into the caller's zeroed HOST arrays of length n. */
long long maze_launch_count(void);
int maze_prof_kernel_count(void);
const char *maze_prof_kernel_name(int kid);
int maze_prof_enable(int on);
int maze_prof_collect(double *ms_host, long long *counts_host, int n);

#ifdef __cplusplus
}
#endif
#endif /* MAZE_B200_H */
