/*
 * maze_oracle.c -- CPU restatement of the LOKI re-segmentation hot path of MAZE-IPP.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import, link or
 * execute this file; it is the checker used by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   - isotropic_* and merge_labels: PINNED against the reference's own files
 *     (maze_ipp/isotropic.py, maze_ipp/merge_labels.py imported from /root/reference)
 *     through the golden vectors in tests/golden/ (tests/golden/make_golden.py).
 *   - label (8-connectivity, raster order): PINNED against scipy.ndimage.label, which is
 *     what skimage.measure.label delegates to for bool input (loki/pipeline.py:430-433).
 *   - clear_border / remove_small_objects / regionprops: PARITY UNPINNED -- scikit-image and
 *     morphocut are not vendored in the reference and not installed; formulas are restated
 *     from the published skimage algorithms and cross-checked against OpenCV moments.
 *
 * Every function cites the reference lines it follows (paths relative to the reference root).
 *
 * Build: gcc -O2 -shared -fPIC -o oracle/libmaze_oracle.so oracle/maze_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_OK 0
#define ORACLE_ERR_TYPEERROR (-2) /* reference raises TypeError (merge_labels.py:19-20) */
#define ORACLE_ERR_NOMEM (-3)

static const int64_t BIG = ((int64_t)1) << 40;

/* ------------------------------------------------------------------------------------------
 * Exact squared Euclidean distance transform.
 * Follows scipy.ndimage.distance_transform_edt as used at maze_ipp/isotropic.py:35,66 and
 * maze_ipp/merge_labels.py:17,22: every nonzero pixel gets the distance to the nearest zero
 * pixel, zero pixels get 0, the image border is not background.  When the input has no zero
 * pixel at all scipy's feature transform behaves as if a single background pixel sat at
 * index (-1, 0) (probed in SURVEY.md section 7 "No-background phantom"); restated here.
 * d2 receives the exact integer squared distance.
 * ---------------------------------------------------------------------------------------- */
void oracle_edt_sq(const uint8_t *img, int H, int W, int64_t *d2)
{
    int64_t *g = (int64_t *)malloc(sizeof(int64_t) * (size_t)H * W); /* vertical distance */
    int any_bg = 0;
    for (int64_t i = 0; i < (int64_t)H * W; i++)
        if (!img[i]) { any_bg = 1; break; }

    for (int x = 0; x < W; x++) {
        /* nearest zero above (or the phantom at row -1 in column 0) */
        int64_t last = -BIG;
        if (!any_bg && x == 0) last = -1;
        for (int y = 0; y < H; y++) {
            if (!img[(int64_t)y * W + x]) last = y;
            g[(int64_t)y * W + x] = (last <= -BIG) ? BIG : (y - last);
        }
        last = BIG;
        for (int y = H - 1; y >= 0; y--) {
            if (!img[(int64_t)y * W + x]) last = y;
            if (last < BIG && last - y < g[(int64_t)y * W + x]) g[(int64_t)y * W + x] = last - y;
        }
    }
    for (int y = 0; y < H; y++) {
        const int64_t *gr = g + (int64_t)y * W;
        for (int x = 0; x < W; x++) {
            int64_t best = (gr[x] >= BIG) ? (BIG * 4) : gr[x] * gr[x];
            /* pruned outward scan: a column k away cannot beat best once k*k >= best */
            for (int64_t k = 1; k * k < best; k++) {
                int found_any = 0;
                if (x - k >= 0) {
                    found_any = 1;
                    int64_t gv = gr[x - k];
                    if (gv < BIG) { int64_t c = k * k + gv * gv; if (c < best) best = c; }
                }
                if (x + k < W) {
                    found_any = 1;
                    int64_t gv = gr[x + k];
                    if (gv < BIG) { int64_t c = k * k + gv * gv; if (c < best) best = c; }
                }
                if (!found_any) break;
            }
            d2[(int64_t)y * W + x] = best;
        }
    }
    free(g);
}

/* Brute force version of the same definition, O(N * #background); for tiny pinning cases only. */
void oracle_edt_sq_bruteforce(const uint8_t *img, int H, int W, int64_t *d2)
{
    int any_bg = 0;
    for (int64_t i = 0; i < (int64_t)H * W; i++)
        if (!img[i]) { any_bg = 1; break; }
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int64_t best = BIG * 4;
            if (!any_bg) {
                best = (int64_t)(y + 1) * (y + 1) + (int64_t)x * x;
            } else {
                for (int v = 0; v < H; v++)
                    for (int u = 0; u < W; u++)
                        if (!img[(int64_t)v * W + u]) {
                            int64_t c = (int64_t)(y - v) * (y - v) + (int64_t)(x - u) * (x - u);
                            if (c < best) best = c;
                        }
            }
            d2[(int64_t)y * W + x] = best;
        }
}

/* maze_ipp/isotropic.py:35-36 -- dist = edt(image); return dist > radius (float64 compare). */
void oracle_isotropic_erosion(const uint8_t *img, int H, int W, double radius, uint8_t *out)
{
    int64_t n = (int64_t)H * W;
    int64_t *d2 = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    oracle_edt_sq(img, H, W, d2);
    for (int64_t i = 0; i < n; i++) out[i] = sqrt((double)d2[i]) > radius;
    free(d2);
}

/* maze_ipp/isotropic.py:66-67 -- dist = edt(image == 0); return dist < radius (STRICT). */
void oracle_isotropic_dilation(const uint8_t *img, int H, int W, double radius, uint8_t *out)
{
    int64_t n = (int64_t)H * W;
    int64_t *d2 = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    uint8_t *inv = (uint8_t *)malloc((size_t)n);
    for (int64_t i = 0; i < n; i++) inv[i] = (img[i] == 0);
    oracle_edt_sq(inv, H, W, d2);
    for (int64_t i = 0; i < n; i++) out[i] = sqrt((double)d2[i]) < radius;
    free(inv);
    free(d2);
}

/* maze_ipp/isotropic.py:97-98 -- erosion then dilation, same radius. */
void oracle_isotropic_opening(const uint8_t *img, int H, int W, double radius, uint8_t *out)
{
    uint8_t *tmp = (uint8_t *)malloc((size_t)H * W);
    oracle_isotropic_erosion(img, H, W, radius, tmp);
    oracle_isotropic_dilation(tmp, H, W, radius, out);
    free(tmp);
}

/* maze_ipp/isotropic.py:128-129 -- dilation then erosion, same radius. */
void oracle_isotropic_closing(const uint8_t *img, int H, int W, double radius, uint8_t *out)
{
    uint8_t *tmp = (uint8_t *)malloc((size_t)H * W);
    oracle_isotropic_dilation(img, H, W, radius, tmp);
    oracle_isotropic_erosion(tmp, H, W, radius, out);
    free(tmp);
}

/* loki/pipeline.py:649 -- mask = image > threshold_brighter (uint8 against a float). */
void oracle_threshold(const uint8_t *img, int64_t n, double thr, uint8_t *out)
{
    for (int64_t i = 0; i < n; i++) out[i] = ((double)img[i] > thr);
}

/* ------------------------------------------------------------------------------------------
 * 8-connected component labelling, labels 1..N in raster order of each component's first
 * pixel, background 0, int32 (loki/pipeline.py:430-433: skimage.measure.label(bool mask) ==
 * scipy.ndimage.label(mask, structure=ones((3,3)))).  Restated as a raster scan with a
 * flood fill per unvisited foreground pixel.  Returns N.
 * ---------------------------------------------------------------------------------------- */
int32_t oracle_label8(const uint8_t *mask, int H, int W, int32_t *labels)
{
    int64_t n = (int64_t)H * W;
    memset(labels, 0, sizeof(int32_t) * (size_t)n);
    int64_t *stack = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    int32_t cur = 0;
    for (int64_t i = 0; i < n; i++) {
        if (!mask[i] || labels[i]) continue;
        cur++;
        int64_t sp = 0;
        stack[sp++] = i;
        labels[i] = cur;
        while (sp) {
            int64_t p = stack[--sp];
            int y = (int)(p / W), x = (int)(p % W);
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    int v = y + dy, u = x + dx;
                    if (v < 0 || v >= H || u < 0 || u >= W) continue;
                    int64_t q = (int64_t)v * W + u;
                    if (mask[q] && !labels[q]) { labels[q] = cur; stack[sp++] = q; }
                }
        }
    }
    free(stack);
    return cur;
}

/* loki/pipeline.py:435-439 -- skimage.segmentation.clear_border(labels, out=labels): every
 * label that has a pixel on the outermost rows/columns is set to 0, in place, no renumbering.
 * (Restated from the published skimage algorithm; the label image comes straight from
 * label(), so skimage's internal re-labelling is the identity partition.) */
void oracle_clear_border(int32_t *labels, int H, int W)
{
    int64_t n = (int64_t)H * W;
    int32_t maxl = 0;
    for (int64_t i = 0; i < n; i++) if (labels[i] > maxl) maxl = labels[i];
    uint8_t *kill = (uint8_t *)calloc((size_t)maxl + 1, 1);
    for (int x = 0; x < W; x++) {
        if (H > 0) { kill[labels[x] > 0 ? labels[x] : 0] = 1; kill[labels[(int64_t)(H - 1) * W + x] > 0 ? labels[(int64_t)(H - 1) * W + x] : 0] = 1; }
    }
    for (int y = 0; y < H; y++) {
        if (W > 0) { int32_t a = labels[(int64_t)y * W], b = labels[(int64_t)y * W + W - 1]; kill[a > 0 ? a : 0] = 1; kill[b > 0 ? b : 0] = 1; }
    }
    kill[0] = 0;
    for (int64_t i = 0; i < n; i++) if (labels[i] > 0 && kill[labels[i]]) labels[i] = 0;
    free(kill);
}

/* loki/pipeline.py:442-448 -- skimage.morphology.remove_small_objects(labels, min_size, out=labels):
 * bincount of the label image, labels with count < min_size are zeroed in place. */
void oracle_remove_small_objects(int32_t *labels, int H, int W, int64_t min_size)
{
    int64_t n = (int64_t)H * W;
    int32_t maxl = 0;
    for (int64_t i = 0; i < n; i++) if (labels[i] > maxl) maxl = labels[i];
    int64_t *cnt = (int64_t *)calloc((size_t)maxl + 1, sizeof(int64_t));
    for (int64_t i = 0; i < n; i++) if (labels[i] > 0) cnt[labels[i]]++;
    for (int64_t i = 0; i < n; i++) if (labels[i] > 0 && cnt[labels[i]] < min_size) labels[i] = 0;
    free(cnt);
}

/* ------------------------------------------------------------------------------------------
 * merge_labels -- maze_ipp/merge_labels.py:29-113 with helpers :7-26.
 * ---------------------------------------------------------------------------------------- */

/* merge_labels.py:12-26.  mask is "labels == l"; writes float64 distances into out.
 * have_max == 0 restates the `max_distance is None` branch (:16-17).  Returns
 * ORACLE_ERR_TYPEERROR where the reference raises (find_objects yields None, :19-20). */
static int windowed_distance_outside(const uint8_t *mask, int H, int W, int have_max, int64_t max_distance_int, double *out)
{
    int64_t n = (int64_t)H * W;
    if (!have_max) {
        uint8_t *inv = (uint8_t *)malloc((size_t)n);
        int64_t *d2 = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
        for (int64_t i = 0; i < n; i++) inv[i] = !mask[i];
        oracle_edt_sq(inv, H, W, d2);
        for (int64_t i = 0; i < n; i++) out[i] = sqrt((double)d2[i]);
        free(inv); free(d2);
        return ORACLE_OK;
    }
    int r0 = H, r1 = -1, c0 = W, c1 = -1;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
            if (mask[(int64_t)y * W + x]) {
                if (y < r0) r0 = y;
                if (y > r1) r1 = y;
                if (x < c0) c0 = x;
                if (x > c1) c1 = x;
            }
    if (r1 < 0) return ORACLE_ERR_TYPEERROR;
    int64_t pad = max_distance_int + 1; /* :20 */
    int64_t wr0 = r0 - pad, wr1 = (int64_t)r1 + 1 + pad, wc0 = c0 - pad, wc1 = (int64_t)c1 + 1 + pad;
    if (wr0 < 0) wr0 = 0;              /* :9  max(0, start - padding) */
    if (wc0 < 0) wc0 = 0;
    if (wr1 > H) wr1 = H;              /* slicing clips the unclipped stop */
    if (wc1 > W) wc1 = W;
    /* a negative stop would be python-style wraparound; cannot happen for pad >= 0 */
    int h = (int)(wr1 - wr0), w = (int)(wc1 - wc0);
    uint8_t *inv = (uint8_t *)malloc((size_t)h * w);
    int64_t *d2 = (int64_t *)malloc(sizeof(int64_t) * (size_t)h * w);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) inv[(int64_t)y * w + x] = !mask[(wr0 + y) * W + wc0 + x];
    oracle_edt_sq(inv, h, w, d2);
    int64_t mx = 0;
    for (int64_t i = 0; i < (int64_t)h * w; i++) if (d2[i] > mx) mx = d2[i];
    double fill = sqrt((double)mx); /* :24 np.full(shape, dist_sliced.max()) */
    for (int64_t i = 0; i < n; i++) out[i] = fill;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) out[(wr0 + y) * W + wc0 + x] = sqrt((double)d2[(int64_t)y * w + x]);
    free(inv); free(d2);
    return ORACLE_OK;
}

static int cmp_i32(const void *a, const void *b)
{
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

/*
 * labels      : int32 H*W, the array the loop keeps READING (merge_labels.py:83, 87, 98).
 * labels_out  : array that receives the writes (:68, 106).  Pass the same pointer as `labels`
 *               to restate the pipeline's aliased call (loki/pipeline.py:452-457).  When the
 *               reference would copy (labels_out=None, :62-63) the caller passes a copy.
 * index_in    : optional ordered list of labels (n_index entries), NULL => sorted unique > 0 (:55-57).
 * have_max    : 0 restates max_distance=None.
 * merge_dists : optional, capacity n_index (or number of labels); *n_merge receives the count.
 * Returns ORACLE_OK, or ORACLE_ERR_TYPEERROR where the reference raises TypeError.
 * *n_index_out receives the number of labels considered (so the caller can restate the
 * `len(index) < 2: return labels` identity return, :59-60).
 */
int oracle_merge_labels(const int32_t *labels, int H, int W, const int32_t *index_in, int n_index,
                        int have_max, double max_distance, double path_tolerance,
                        int32_t *labels_out, double *merge_dists, int *n_merge, int *n_index_out)
{
    int64_t n = (int64_t)H * W;
    int32_t *index = NULL;
    int ni = 0;
    if (n_merge) *n_merge = 0;
    if (index_in) {
        ni = n_index;
        index = (int32_t *)malloc(sizeof(int32_t) * (size_t)(ni > 0 ? ni : 1));
        memcpy(index, index_in, sizeof(int32_t) * (size_t)ni);
    } else {
        int32_t *tmp = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
        int64_t m = 0;
        for (int64_t i = 0; i < n; i++) if (labels[i] > 0) tmp[m++] = labels[i];
        qsort(tmp, (size_t)m, sizeof(int32_t), cmp_i32);
        for (int64_t i = 0; i < m; i++) if (i == 0 || tmp[i] != tmp[i - 1]) tmp[ni++] = tmp[i];
        index = tmp;
    }
    if (n_index_out) *n_index_out = ni;
    if (ni < 2) { free(index); return ORACLE_OK; } /* :59-60 */

    uint8_t *mask = (uint8_t *)malloc((size_t)n);
    double *distmap = (double *)malloc(sizeof(double) * (size_t)n);
    double *cur = (double *)malloc(sizeof(double) * (size_t)n);
    int rc = ORACLE_OK;
    int64_t max_distance_int = have_max ? (int64_t)ceil(max_distance) : 0; /* :70 */

    int32_t l0 = index[0]; /* :66 */
    memmove(index, index + 1, sizeof(int32_t) * (size_t)(ni - 1));
    ni--;
    for (int64_t i = 0; i < n; i++) mask[i] = (labels[i] == l0);
    for (int64_t i = 0; i < n; i++) if (mask[i]) labels_out[i] = l0; /* :68 */

    rc = windowed_distance_outside(mask, H, W, have_max, max_distance_int, distmap); /* :73 */
    if (rc) goto done;
    double max_dist = 0;
    for (int64_t i = 0; i < n; i++) if (distmap[i] > max_dist) max_dist = distmap[i]; /* :74 */
    /* labelmap (:77) only ever holds l0, so replace_label (:103) is always l0. */

    while (ni > 0) { /* :81 */
        int best_k = 0;
        double best_v = 0;
        for (int k = 0; k < ni; k++) { /* :83, np.argmin keeps the first minimum */
            double v = max_dist;
            int32_t l = index[k];
            for (int64_t i = 0; i < n; i++) if (labels[i] == l && distmap[i] < v) v = distmap[i];
            if (k == 0 || v < best_v) { best_v = v; best_k = k; }
        }
        int32_t cur_l = index[best_k]; /* :84 */
        memmove(index + best_k, index + best_k + 1, sizeof(int32_t) * (size_t)(ni - best_k - 1));
        ni--;

        for (int64_t i = 0; i < n; i++) mask[i] = (labels[i] == cur_l);
        rc = windowed_distance_outside(mask, H, W, have_max, max_distance_int, cur); /* :87 */
        if (rc) goto done;

        double merge_dist = INFINITY;
        for (int64_t i = 0; i < n; i++) { double s = distmap[i] + cur[i]; if (s < merge_dist) merge_dist = s; } /* :90-92 */
        if (have_max && merge_dist > max_distance) break; /* :94-96 */

        double lim = merge_dist + path_tolerance;
        if (merge_dists && n_merge) merge_dists[*n_merge] = merge_dist; /* :100 */
        if (n_merge) (*n_merge)++;
        /* :98 and :106.  mask was evaluated from `labels` before any write of this iteration,
         * exactly as numpy evaluates the right-hand side before the fancy assignment. */
        for (int64_t i = 0; i < n; i++)
            if (mask[i] || (distmap[i] + cur[i] <= lim)) labels_out[i] = l0;
        for (int64_t i = 0; i < n; i++) if (cur[i] < distmap[i]) distmap[i] = cur[i]; /* :109-111 */
    }
done:
    free(mask); free(distmap); free(cur); free(index);
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * regionprops subset -- what loki/pipeline.py:604-625, 653-654 read through skimage's
 * RegionProperties (PARITY UNPINNED, see header).  One row of ORACLE_NFEAT doubles per label
 * 1..max_label; rows of absent labels have area 0 and the rest NaN.
 * Column layout is shared with the product (include/maze_b200.h, MAZE_F_*).
 * ---------------------------------------------------------------------------------------- */
#define ORACLE_NFEAT 64
enum {
    F_LABEL = 0, F_AREA = 1, F_BBOX = 2 /*4*/, F_CENTROID = 6 /*2*/, F_MU = 8 /*16, mu[p*4+q]*/,
    F_NU = 24 /*16*/, F_HU = 40 /*7*/, F_EIG = 47 /*2*/, F_AXIS_MAJOR = 49, F_AXIS_MINOR = 50,
    F_ECC = 51, F_ORIENT = 52, F_IMIN = 53, F_IMAX = 54, F_IMEAN = 55, F_FRAC_INVALID = 56,
    F_IMAGE = 57, F_T00 = 58, F_T01 = 59, F_T11 = 60
};

int oracle_regionprops(const int32_t *labels, const uint8_t *image, int H, int W, int32_t max_label, double *table)
{
    const double PI = 3.14159265358979323846;
    int64_t nl = max_label;
    for (int64_t i = 0; i < nl * ORACLE_NFEAT; i++) table[i] = NAN;
    if (nl <= 0) return ORACLE_OK;
    int64_t *area = (int64_t *)calloc((size_t)nl + 1, sizeof(int64_t));
    int64_t *sr = (int64_t *)calloc((size_t)nl + 1, sizeof(int64_t));
    int64_t *sc = (int64_t *)calloc((size_t)nl + 1, sizeof(int64_t));
    int64_t *si = (int64_t *)calloc((size_t)nl + 1, sizeof(int64_t));
    int64_t *sz = (int64_t *)calloc((size_t)nl + 1, sizeof(int64_t));
    int *r0 = (int *)malloc(sizeof(int) * ((size_t)nl + 1)), *r1 = (int *)malloc(sizeof(int) * ((size_t)nl + 1));
    int *c0 = (int *)malloc(sizeof(int) * ((size_t)nl + 1)), *c1 = (int *)malloc(sizeof(int) * ((size_t)nl + 1));
    int *imn = (int *)malloc(sizeof(int) * ((size_t)nl + 1)), *imx = (int *)malloc(sizeof(int) * ((size_t)nl + 1));
    for (int64_t l = 0; l <= nl; l++) { r0[l] = H; c0[l] = W; r1[l] = -1; c1[l] = -1; imn[l] = 256; imx[l] = -1; }
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int32_t l = labels[(int64_t)y * W + x];
            if (l <= 0 || l > nl) continue;
            int v = image ? image[(int64_t)y * W + x] : 0;
            area[l]++; sr[l] += y; sc[l] += x; si[l] += v; sz[l] += (v == 0);
            if (y < r0[l]) r0[l] = y;
            if (y > r1[l]) r1[l] = y;
            if (x < c0[l]) c0[l] = x;
            if (x > c1[l]) c1[l] = x;
            if (v < imn[l]) imn[l] = v;
            if (v > imx[l]) imx[l] = v;
        }
    for (int64_t l = 1; l <= nl; l++) {
        double *f = table + (l - 1) * ORACLE_NFEAT;
        f[F_LABEL] = (double)l;
        f[F_AREA] = (double)area[l];
        f[F_IMAGE] = 0;
        if (!area[l]) continue;
        f[F_BBOX + 0] = r0[l]; f[F_BBOX + 1] = c0[l]; f[F_BBOX + 2] = r1[l] + 1; f[F_BBOX + 3] = c1[l] + 1;
        double cr = (double)sr[l] / (double)area[l], cc = (double)sc[l] / (double)area[l];
        f[F_CENTROID] = cr; f[F_CENTROID + 1] = cc;
        /* central moments about the centroid, bbox-local like skimage's moments_central */
        double mu[4][4] = {{0}};
        double lr = cr - r0[l], lc = cc - c0[l];
        for (int y = r0[l]; y <= r1[l]; y++) {
            double rowq[4] = {0, 0, 0, 0};
            for (int x = c0[l]; x <= c1[l]; x++)
                if (labels[(int64_t)y * W + x] == l) {
                    double dc = (double)(x - c0[l]) - lc, p = 1;
                    for (int q = 0; q < 4; q++) { rowq[q] += p; p *= dc; }
                }
            double dr = (double)(y - r0[l]) - lr, pr = 1;
            for (int p = 0; p < 4; p++) { for (int q = 0; q < 4; q++) mu[p][q] += pr * rowq[q]; pr *= dr; }
        }
        for (int p = 0; p < 4; p++) for (int q = 0; q < 4; q++) f[F_MU + p * 4 + q] = mu[p][q];
        double nu[4][4];
        for (int p = 0; p < 4; p++)
            for (int q = 0; q < 4; q++) {
                nu[p][q] = (p + q >= 2) ? mu[p][q] / pow(mu[0][0], (p + q) / 2.0 + 1.0) : NAN;
                f[F_NU + p * 4 + q] = nu[p][q];
            }
        { /* Hu invariants, nu indexed [row power][col power] */
            double t0 = nu[3][0] + nu[1][2], t1 = nu[2][1] + nu[0][3];
            double q0 = t0 * t0, q1 = t1 * t1;
            double n4 = 4 * nu[1][1], s = nu[2][0] + nu[0][2], d = nu[2][0] - nu[0][2];
            double *hu = f + F_HU;
            hu[0] = s;
            hu[1] = d * d + n4 * nu[1][1];
            hu[3] = q0 + q1;
            hu[5] = d * (q0 - q1) + n4 * t0 * t1;
            t0 *= q0 - 3 * q1;
            t1 *= 3 * q0 - q1;
            q0 = nu[3][0] - 3 * nu[1][2];
            q1 = 3 * nu[2][1] - nu[0][3];
            hu[2] = q0 * q0 + q1 * q1;
            hu[4] = q0 * t0 + q1 * t1;
            hu[6] = q1 * t0 - q0 * t1;
        }
        double a = mu[0][2] / mu[0][0], b = -mu[1][1] / mu[0][0], c = mu[2][0] / mu[0][0];
        f[F_T00] = a; f[F_T01] = b; f[F_T11] = c;
        double tr = 0.5 * (a + c), df = 0.5 * (a - c);
        double rad = sqrt(df * df + b * b);
        double l1 = tr + rad, l2 = tr - rad;
        if (l1 < 0) l1 = 0;
        if (l2 < 0) l2 = 0;
        f[F_EIG] = l1; f[F_EIG + 1] = l2;
        f[F_AXIS_MAJOR] = 4 * sqrt(l1);
        f[F_AXIS_MINOR] = 4 * sqrt(l2);
        f[F_ECC] = (l1 == 0) ? 0.0 : sqrt(1 - l2 / l1);
        if (a - c == 0) f[F_ORIENT] = (b < 0) ? PI / 4 : -PI / 4;
        else f[F_ORIENT] = 0.5 * atan2(-2 * b, c - a);
        if (image) {
            f[F_IMIN] = imn[l]; f[F_IMAX] = imx[l];
            f[F_IMEAN] = (double)si[l] / (double)area[l];
            f[F_FRAC_INVALID] = (double)sz[l] / (double)area[l]; /* loki/pipeline.py:617 */
        }
    }
    free(area); free(sr); free(sc); free(si); free(sz);
    free(r0); free(r1); free(c0); free(c1); free(imn); free(imx);
    return ORACLE_OK;
}

int oracle_nfeat(void) { return ORACLE_NFEAT; }
