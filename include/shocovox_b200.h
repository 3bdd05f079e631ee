/*
 * shocovox_b200 — C ABI of the B200-native primary-ray path of davids91/shocovox.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / C++ types. Every entry point cites
 * the reference (shocovox-rs 0.11.1) interface it replaces; paths are relative to the reference checkout.
 * The reference has no FFI of its own (it is a Rust crate), so these are the symbols a `shocovox_b200-sys`
 * style `extern "C"` block binds (INTEGRATION.md shows that block).
 *
 * Rules
 *  - every call returns an svx_status (0 = OK); nothing throws or aborts across the boundary;
 *  - handles are created by the library and released by the caller with the matching *_free, in ANY order: a host keeps
 *    its octree alive and a view keeps its host alive until they are freed themselves (bindings whose finalisers run in
 *    no particular order are safe); a freed handle must not be used again;
 *  - an octree handle allows one writer or many readers (caller-enforced, like `&mut self` / `&self`);
 *  - a host/view handle serialises its own calls internally and owns its CUDA stream;
 *  - ray queries run on the GPU only. There is no CPU fallback: without a CUDA device every svx_gpu_* /
 *    svx_view_* call fails with SVX_E_CUDA.
 */
#ifndef SHOCOVOX_B200_H
#define SHOCOVOX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVX_API __attribute__((visibility("default")))

/* OctreeError, src/octree/types.rs:9-21 (+ device-side failures) */
typedef enum svx_status {
    SVX_OK = 0,
    SVX_E_INVALID_SIZE = 1,            /* OctreeError::InvalidSize            */
    SVX_E_INVALID_BRICK_DIMENSION = 2, /* OctreeError::InvalidBrickDimension  */
    SVX_E_INVALID_STRUCTURE = 3,       /* OctreeError::InvalidStructure       */
    SVX_E_INVALID_POSITION = 4,        /* OctreeError::InvalidPosition        */
    SVX_E_INVALID_ARGUMENT = 5,        /* null handle / bad enum / zero resolution */
    SVX_E_DECODE = 6,                  /* from_bytes / load: not a bencoded Octree (the reference panics, octree/mod.rs:147) */
    SVX_E_IO = 7,                      /* save / load: std::io::Error */
    SVX_E_TIMEOUT = 8,                 /* multi-GPU gather: a member did not deliver its share of the frame in time */
    SVX_E_CUDA = -1,                   /* CUDA runtime error (svx_last_error_message has the text) */
    SVX_E_OUT_OF_MEMORY = -2
} svx_status;

/* Albedo, src/octree/types.rs:92-97 ; Albedo::from(u32) is 0xRRGGBBAA, src/octree/detail.rs:92-105 */
typedef struct svx_albedo {
    uint8_t r, g, b, a;
} svx_albedo;

/* OctreeEntry<u32>, src/octree/types.rs:24-36 */
typedef enum svx_entry_kind {
    SVX_ENTRY_EMPTY = 0,
    SVX_ENTRY_VISUAL = 1,
    SVX_ENTRY_INFORMATIVE = 2,
    SVX_ENTRY_COMPLEX = 3
} svx_entry_kind;

typedef struct svx_entry {
    uint32_t kind; /* svx_entry_kind */
    svx_albedo albedo;
    uint32_t data;
} svx_entry;

/* Ray, src/spatial/raytracing/mod.rs:8-11 ; direction must be unit length (:14-16) */
typedef struct svx_ray {
    float origin[3];
    float direction[3];
} svx_ray;

/* Return tuple of get_by_ray, src/raytracing/raytracing_on_cpu.rs:316 :
 * Option<(OctreeEntry<T>, V3c<f32> impact_point, V3c<f32> normal)> plus the derived outputs of SURVEY §8 R14 */
typedef struct svx_hit {
    uint32_t hit;           /* 0 = None */
    uint32_t palette_value; /* PaletteIndexValues of the voxel (colour idx | data idx << 16); 0xFFFFFFFF on a miss */
    svx_entry entry;        /* palette-resolved entry */
    float impact_point[3];
    float normal[3];
    float distance;         /* (impact_point - origin).length() */
} svx_hit;

/* Viewport, src/raytracing/bevy/types.rs:55-71 */
typedef struct svx_viewport {
    float origin[3];
    float direction[3];
    float frustum[3]; /* x: glass width, y: glass height, z: max depth */
    float fov;
} svx_viewport;

/* The examples disagree on where the looking glass sits: `direction * fov`
 * (examples/cpu_render.rs:92, beach.rs:180, minecraft.rs:177) or `direction * frustum.z`
 * (examples/dot_cube.rs:209, sponza.rs:179). The view takes it as a mode. */
typedef enum svx_glass_mode { SVX_GLASS_AT_FOV = 0, SVX_GLASS_AT_FRUSTUM_Z = 1 } svx_glass_mode;

/* What the GPUs of a tile-sharded frame send to the GPU that assembles it (svx_view_gather_*, svx_multi_*) */
typedef enum svx_wire_format {
    SVX_WIRE_THREE_PLANES = 0, /* hit id, albedo and distance: 12 bytes per pixel cross NVLink */
    SVX_WIRE_ID_DISTANCE = 1   /* hit id and distance, 8 bytes per pixel; the assembling GPU resolves albedo =
                                  palette[hit_id & 0xFFFF] for the received rows (same tree, same palette: same bytes) */
} svx_wire_format;

/* Everything another process needs to render into a root view's framebuffer; plain bytes, ship them any way you like */
typedef struct svx_gather_handle {
    uint8_t ipc[64];         /* cudaIpcMemHandle_t of the root's frame allocation (three planes + completion flags) */
    uint32_t width, height;  /* the root's resolution: a joining view must have the same */
    uint32_t world, rows_per_band;
    int32_t wire;            /* svx_wire_format */
    int32_t device;          /* CUDA device ordinal of the root */
    uint64_t plane_bytes;    /* distance between the planes inside the allocation */
    uint64_t generation;     /* changes when the root reallocates its frame */
    uint8_t reserved[24];
} svx_gather_handle;

typedef struct svx_octree svx_octree;     /* Octree<u32>,    src/octree/types.rs:169-207      */
typedef struct svx_gpu_host svx_gpu_host; /* OctreeGPUHost,  src/raytracing/bevy/types.rs:80-87 */
typedef struct svx_view svx_view;         /* OctreeGPUView,  src/raytracing/bevy/types.rs:92-130 */
typedef struct svx_multi svx_multi;       /* one OctreeGPUHost + OctreeGPUView per GPU of this process, rendering one frame together */

/* Device-resident frame produced by svx_view_render: SoA, image order (row 0 = top, pixel (x, y) of the
 * reference's caller loop lands in row h-1-y, examples/cpu_render.rs:106). Pointers are CUDA device pointers
 * owned by the view and valid until the next render / set_resolution / free on that view. */
typedef struct svx_frame {
    uint32_t width, height;
    uint32_t row_begin, row_end; /* rows this view renders (its shard); others are left untouched */
    const uint32_t* hit_id;      /* [h*w] palette value of the hit voxel, 0xFFFFFFFF = miss */
    const uint32_t* albedo;      /* [h*w] RGBA8 (r in the low byte), 0 on a miss or when the voxel has no colour */
    const float* distance;       /* [h*w] hit distance, 0 on a miss */
    float kernel_ms;             /* CUDA-event time of the traversal kernel for this render */
} svx_frame;

/* Counters of one serialised tree (render-data upload) */
typedef struct svx_gpu_stats {
    uint64_t nodes, bricks, voxel_bytes, total_bytes;
    uint32_t tree_size, brick_dim, depth, colours;
} svx_gpu_stats;

/* What the last render-data upload of a host moved: everything the first time, afterwards the node tables plus the
 * bricks written since the previous upload */
typedef struct svx_upload_stats {
    uint64_t bricks; /* bricks copied host -> device */
    uint64_t bytes;  /* bytes copied host -> device (nodes, palettes, voxels, brick handle list) */
    uint32_t full;   /* 1: first upload of this host */
    float bits_kernel_ms; /* CUDA-event time of the kernel that derives the occupancy bit-bricks of the copied bricks */
} svx_upload_stats;

/* ---- library ------------------------------------------------------------------------------------------- */
SVX_API const char* svx_version(void);
SVX_API const char* svx_last_error_message(void); /* thread-local text of the last failing call */
SVX_API int32_t svx_cuda_device_count(void);      /* 0 when no usable CUDA device exists */

/* ---- Octree: construction and point queries (host) --------------------------------------------------- */
/* Octree::new, src/octree/mod.rs:173-205 (validation order kept: brick dimension, size, structure) */
SVX_API int32_t svx_octree_new(uint32_t size, uint32_t brick_dim, svx_octree** out);
SVX_API void svx_octree_free(svx_octree* tree);
/* Octree::insert, src/octree/update/insert.rs:47-56 ; an empty entry is a no-op returning OK (:117-119) */
SVX_API int32_t svx_octree_insert(svx_octree* tree, uint32_t x, uint32_t y, uint32_t z, const svx_entry* entry);
/* Octree::insert_at_lod, src/octree/update/insert.rs:63-73 */
SVX_API int32_t svx_octree_insert_at_lod(svx_octree* tree, uint32_t x, uint32_t y, uint32_t z, uint32_t insert_size,
                                         const svx_entry* entry);
/* Octree::update, src/octree/update/insert.rs:79-88 */
SVX_API int32_t svx_octree_update(svx_octree* tree, uint32_t x, uint32_t y, uint32_t z, const svx_entry* entry);
/* Octree::clear / clear_at_lod, src/octree/update/clear.rs:48-78 */
SVX_API int32_t svx_octree_clear(svx_octree* tree, uint32_t x, uint32_t y, uint32_t z);
SVX_API int32_t svx_octree_clear_at_lod(svx_octree* tree, uint32_t x, uint32_t y, uint32_t z, uint32_t clear_size);
/* The per-voxel insert loop every example runs (examples/cpu_render.rs:21-43): n Visual inserts in array order.
 * xyz is [n][3], rgba is [n][4], lod (optional) holds an insert_at_lod size per voxel (<= 1 means insert). */
SVX_API int32_t svx_octree_insert_batch(svx_octree* tree, const uint32_t* xyz, const uint8_t* rgba, const uint32_t* lod,
                                        uint64_t n);
/* Octree::get, src/octree/mod.rs:209-215 */
SVX_API int32_t svx_octree_get(const svx_octree* tree, uint32_t x, uint32_t y, uint32_t z, svx_entry* out);
/* get() over the box [x0,x0+nx) x [y0,y0+ny) x [z0,z0+nz), x-major then y then z */
SVX_API int32_t svx_octree_get_sweep(const svx_octree* tree, uint32_t x0, uint32_t y0, uint32_t z0, uint32_t nx,
                                     uint32_t ny, uint32_t nz, svx_entry* out);
/* Octree::get_size, src/octree/mod.rs:374-376 */
SVX_API uint32_t svx_octree_size(const svx_octree* tree);
SVX_API uint32_t svx_octree_brick_dim(const svx_octree* tree);
/* pub auto_simplify, src/octree/types.rs:203 */
SVX_API int32_t svx_octree_set_auto_simplify(svx_octree* tree, int32_t enabled);
/* Key-order independent digest of the reachable tree (node kinds, occupancy bits, bricks, palettes) */
SVX_API uint64_t svx_octree_structure_hash(const svx_octree* tree);
SVX_API uint64_t svx_octree_node_count(const svx_octree* tree);
/* The palettes the tree's voxels index (voxel_color_palette / voxel_data_palette, src/octree/types.rs:191-192): a voxel is
 * PaletteIndexValues = colour index | data index << 16, 0xFFFF = none (src/octree/types.rs:100, detail.rs:31-60) - which is
 * also what the hit_id plane of a frame holds. Copies min(capacity, *count) entries; *count receives the palette's size
 * (out may be NULL to ask for it). With these a caller of svx_view_render_to_host(view, hit_id, NULL, distance) resolves
 * albedo = colours[hit_id & 0xFFFF] itself, like the reference's shader does from its color_palette buffer
 * (assets/shaders/viewport_render.wgsl). MIP maps add colours: read the palette after switching them on. */
SVX_API int32_t svx_octree_color_palette(const svx_octree* tree, svx_albedo* out, uint32_t capacity, uint32_t* count);
SVX_API int32_t svx_octree_data_palette(const svx_octree* tree, uint32_t* out, uint32_t capacity, uint32_t* count);

/* ---- MIP maps: Octree::albedo_mip_map_resampling_strategy() -> StrategyUpdater, src/octree/mod.rs:379,
 * src/octree/mipmap.rs:716-938. Every node owns one MIP brick (types.rs:186) holding a simplified view of its content;
 * edits refresh the affected MIP voxels (insert.rs:371, clear.rs:335) and get_by_ray_at_lod reads them. */
typedef enum svx_mip_method { /* MIPResamplingMethods, src/octree/types.rs:106-139 */
    SVX_MIP_BOX_FILTER = 0,      /* gamma-2 average of the cell one level below; adds colours to the palette (default) */
    SVX_MIP_POINT_FILTER = 1,    /* most frequent colour of the cell one level below */
    SVX_MIP_POINT_FILTER_BD = 2, /* most frequent colour of the voxels at the bottom ("bottom dominant") */
    SVX_MIP_POSTERIZE = 3,       /* average of the largest group of similar colours; parameter = similarity in [0, 1] */
    SVX_MIP_POSTERIZE_BD = 4
} svx_mip_method;
/* StrategyUpdater::switch_albedo_mip_maps, mipmap.rs:858-872: enabling recalculates every MIP of a non-empty tree */
SVX_API int32_t svx_octree_switch_albedo_mip_maps(svx_octree* tree, int32_t enabled);
SVX_API int32_t svx_octree_mip_maps_enabled(const svx_octree* tree); /* MIPMapStrategy::is_enabled, mipmap.rs:693 */
/* StrategyUpdater::recalculate_mips, mipmap.rs:798-855 */
SVX_API int32_t svx_octree_recalculate_mips(svx_octree* tree);
/* set_method_at / get_method_at, mipmap.rs:650-672, :758-771 (threshold clamped to [0, 1]; default BoxFilter) */
SVX_API int32_t svx_octree_mip_set_method_at(svx_octree* tree, uint64_t mip_level, int32_t method, float threshold);
SVX_API int32_t svx_octree_mip_get_method_at(const svx_octree* tree, uint64_t mip_level, int32_t* method, float* threshold);
/* set_color_similarity_thr_at / get_new_color_similarity_at, mipmap.rs:610-640, :726-745 */
SVX_API int32_t svx_octree_mip_set_color_similarity_thr_at(svx_octree* tree, uint64_t mip_level, float threshold);
SVX_API float svx_octree_mip_get_color_similarity_at(const svx_octree* tree, uint64_t mip_level);
/* StrategyUpdater::reset, mipmap.rs:718-721: back to MIPMapStrategy::default() (disabled) */
SVX_API int32_t svx_octree_mip_reset(svx_octree* tree);
/* sample_root_mip (mipmap.rs:897-937, the reference's test hook): the MIP voxel (x, y, z) of the root (octant 8) or of
 * the root's child in `octant` */
SVX_API int32_t svx_octree_mip_sample_root(const svx_octree* tree, uint32_t octant, uint32_t x, uint32_t y, uint32_t z,
                                           svx_entry* out);
/* Digest of the strategy and of the MIP bricks of all reachable nodes */
SVX_API uint64_t svx_octree_mip_hash(const svx_octree* tree);

/* Octree::to_bytes / from_bytes / save / load, src/octree/mod.rs:138-168: the bencode byte format of
 * src/convert/bytecode.rs (a tree saved by the Rust crate loads here and the other way round, MIP bricks and MIP
 * strategy included). to_bytes hands out a library-owned buffer: release it with
 * svx_bytes_free. from_bytes / load validate size and brick_dim like Octree::new and return SVX_E_DECODE on
 * malformed input, SVX_E_IO when the file cannot be read or written. */
SVX_API int32_t svx_octree_to_bytes(const svx_octree* tree, uint8_t** bytes, uint64_t* len);
SVX_API void svx_bytes_free(uint8_t* bytes);
SVX_API int32_t svx_octree_from_bytes(const uint8_t* bytes, uint64_t len, svx_octree** out);
SVX_API int32_t svx_octree_save(const svx_octree* tree, const char* path);
SVX_API int32_t svx_octree_load(const char* path, svx_octree** out);

/* MagicaVoxel `.vox` import. Octree::load_vox_file(filename, brick_dimension), src/convert/magicavoxel.rs:266-289: the
 * scene graph of frame 0 is walked (iterate_vox_tree, :105-197), models are placed at translation - size/2 with their
 * 90-degree rotations (:349-385), converted from MagicaVoxel's right-handed z-up to the crate's left-handed y-up (swap of
 * y and z), shifted to the minimum corner, and every voxel is inserted as a Visual entry with its palette colour; the
 * tree size is the next power of two of the largest extent (:266-271). SVX_E_DECODE for a malformed file, for a file
 * without RGBA chunk (MagicaVoxel's built-in default palette is not reproduced) or without a scene graph (the reference
 * panics); SVX_E_INVALID_* when Octree::new refuses (size, brick_dim) - the reference panics there as well.
 * MIPMapStrategy::load_vox_file (:207-250) = svx_vox_required_tree_size + svx_octree_new + the svx_octree_mip_* settings +
 * svx_octree_insert_vox: the strategy is installed on the empty tree, every insert then refreshes the MIPs as it goes. */
SVX_API int32_t svx_octree_load_vox(const char* path, uint32_t brick_dim, svx_octree** out);
SVX_API int32_t svx_octree_load_vox_bytes(const uint8_t* bytes, uint64_t len, uint32_t brick_dim, svx_octree** out);
SVX_API int32_t svx_vox_required_tree_size(const uint8_t* bytes, uint64_t len, uint32_t* tree_size);
SVX_API int32_t svx_octree_insert_vox(svx_octree* tree, const uint8_t* bytes, uint64_t len); /* load_vox_data_internal, :349-385 */

/* ---- OctreeGPUHost: render-data upload ---------------------------------------------------------------- */
/* OctreeGPUHost{tree}, src/raytracing/bevy/types.rs:80-87. Serialises the WHOLE tree into coalesced SoA
 * buffers and uploads it to `device` (the reference streams nodes on demand instead, bevy/data.rs:365). The
 * octree handle must outlive the host. */
SVX_API int32_t svx_gpu_host_create(const svx_octree* tree, int32_t device, svx_gpu_host** out);
SVX_API void svx_gpu_host_free(svx_gpu_host* host);
/* Bring the device copy up to date after the tree was edited (OctreeGPUView::reload, src/raytracing/bevy/mod.rs:56-60;
 * the job of write_to_gpu's per-request node / brick uploads, bevy/data.rs:365-773). Incremental: the node tables and
 * the palette are replaced, of the bricks only those written since this host's last upload are copied. A no-op when
 * the tree has not been modified. Waits for the device first: no render may be in flight on another thread. */
SVX_API int32_t svx_gpu_host_reload(svx_gpu_host* host);
/* Host image of the node table a host uploads: one 64-byte record per node in breadth-first order (root = record 0):
 *   u32[0..1] stored occupied bits lo / hi (src/octree/detail.rs:524-544)
 *   u32[2]    meta: node kind [1:0] (0 Nothing, 1 Internal, 2 Leaf, 3 UniformLeaf), brick kind of octant o at
 *             [3+2o : 2+2o] (0 empty, 1 parted, 2 solid), MIP brick kind [19:18], octant in the parent [22:20]
 *   u32[3]    index of the parent record (0xFFFFFFFF for the root)
 *   u32[4..11] Internal: child record per octant or 0xFFFFFFFF; Leaf: brick slot per octant; UniformLeaf: slot 0
 *   f32[12..15] bounds: min x, y, z, size
 * This is the flattening the reference does in OctreeRenderData (src/raytracing/bevy/types.rs:216-279) for its own shader;
 * it needs no GPU. `records` may be NULL to query the count; otherwise `capacity` records are available and the call
 * fails with SVX_E_INVALID_ARGUMENT when the tree has more. */
SVX_API int32_t svx_octree_render_data_nodes(const svx_octree* tree, void* records, uint64_t capacity, uint64_t* n_nodes);
/* The same plus the second per-node table of the render data: the brick slot of every node's MIP brick (reference
 * node_mips[key], src/octree/types.rs:186; 0xFFFFFFFF = none or MIP maps disabled), read only by the level-of-detail branch
 * of get_by_ray_at_lod (raytracing_on_cpu.rs:368-386). Either output may be NULL. */
SVX_API int32_t svx_octree_render_data_nodes_with_mips(const svx_octree* tree, void* records, uint32_t* mip_slots, uint64_t capacity,
                                                       uint64_t* n_nodes);
/* Host image of the brick part of the render data: the pooled voxel array (`voxels_per_brick` palette values per brick,
 * flat_projection order, indexed by the brick slots of the node records) and its 1-bit-per-voxel occupancy
 * (`words_per_brick` u32 per brick; bit set unless pix_points_to_empty, src/octree/node.rs:405-427 - the device derives the
 * same words itself during an upload). Either buffer may be NULL; both NULL queries the sizes. No GPU needed. */
SVX_API int32_t svx_octree_render_data_bricks(const svx_octree* tree, uint32_t* voxels, uint32_t* bits, uint64_t capacity_bricks,
                                              uint64_t* n_bricks, uint32_t* voxels_per_brick, uint32_t* words_per_brick);
/* RAY_TO_NODE_OCCUPANCY_BITMASK_LUT (src/spatial/lut.rs, regenerated from its generator logic :39-89) in the layout the
 * kernels read: [direction octant 0..8][cell 0..64] {lo, hi} = 1024 u32 */
SVX_API int32_t svx_render_data_ray_lut(uint32_t* lut);
/* What the most recent upload / reload of this host copied */
SVX_API int32_t svx_gpu_host_last_upload(const svx_gpu_host* host, svx_upload_stats* out);
SVX_API int32_t svx_gpu_host_stats(const svx_gpu_host* host, svx_gpu_stats* out);
/* Octree::get_by_ray (src/raytracing/raytracing_on_cpu.rs:316-318) for n rays at once, on the GPU.
 * `rays` and `hits` are HOST arrays; n = 1 is the reference's single-ray call. */
SVX_API int32_t svx_gpu_host_get_by_rays(svx_gpu_host* host, const svx_ray* rays, uint64_t n, svx_hit* hits);
/* Octree::get_by_ray_at_lod (src/raytracing/raytracing_on_cpu.rs:325-565): the same query with the level-of-detail
 * branch (:368-386) - a node whose MIP level is below `travelled distance / viewing_distance` answers with its MIP
 * brick. viewing_distance = FLT_MAX is get_by_ray. MIP hits carry no user data (the reference's own warning). */
SVX_API int32_t svx_gpu_host_get_by_rays_at_lod(svx_gpu_host* host, const svx_ray* rays, uint64_t n, float viewing_distance,
                                                svx_hit* hits);

/* Octree::get_by_ray(&Ray) / get_by_ray_at_lod(&Ray, viewing_distance) on the tree handle itself
 * (src/raytracing/raytracing_on_cpu.rs:316-325): one ray, answered on the GPU. The tree keeps a device copy on device 0,
 * created on first use and brought up to date before every query; without a CUDA device the call fails with SVX_E_CUDA
 * (there is no CPU ray path). For many rays use svx_gpu_host_get_by_rays. */
SVX_API int32_t svx_octree_get_by_ray(svx_octree* tree, const svx_ray* ray, svx_hit* out);
SVX_API int32_t svx_octree_get_by_ray_at_lod(svx_octree* tree, const svx_ray* ray, float viewing_distance, svx_hit* out);

/* ---- OctreeGPUView: viewport + framebuffer -------------------------------------------------------------- */
/* OctreeGPUHost::create_new_view, src/raytracing/bevy/data.rs:111-166. `size_hint` is the reference's node-cache
 * capacity; the whole tree is resident here so it is ignored. resolution = [width, height]. */
SVX_API int32_t svx_gpu_host_create_view(svx_gpu_host* host, uint32_t size_hint, const svx_viewport* viewport,
                                         uint32_t width, uint32_t height, svx_view** out);
SVX_API void svx_view_free(svx_view* view);
/* OctreeGPUView::reload, src/raytracing/bevy/mod.rs:56-60: svx_gpu_host_reload of the host this view renders */
SVX_API int32_t svx_view_reload(svx_view* view);
/* OctreeSpyGlass::viewport / viewport_mut, src/raytracing/bevy/mod.rs:90-99 */
SVX_API int32_t svx_view_get_viewport(const svx_view* view, svx_viewport* out);
SVX_API int32_t svx_view_set_viewport(svx_view* view, const svx_viewport* viewport);
SVX_API int32_t svx_view_set_glass_mode(svx_view* view, int32_t mode /* svx_glass_mode */);
/* The viewing distance every pixel's get_by_ray_at_lod uses. Default FLT_MAX = get_by_ray (no level of detail); the
 * reference's own GPU path feeds viewport.frustum.z here (assets/shaders/viewport_render.wgsl:631-638). Ignored while
 * the tree's MIP maps are disabled (raytracing_on_cpu.rs:369). */
SVX_API int32_t svx_view_set_viewing_distance(svx_view* view, float viewing_distance);
SVX_API int32_t svx_view_get_viewing_distance(const svx_view* view, float* viewing_distance);
/* Optional fourth framebuffer plane: the pixel the reference's caller loops write (examples/cpu_render.rs:119-136,
 * examples/dot_cube.rs:238-256): albedo.rgb scaled by `1 - (normal . light / 2 + 0.5)`, each channel `as u8`, grey
 * (128,128,128) on a miss, alpha 255; RGBA8 with r in the low byte. light_normal[3] is the examples'
 * `diffuse_light_normal` (they use normalized(0,-1,1)); null switches the plane off again. While it is on,
 * svx_view_render / svx_view_render_to_host also fill it (static schedule; the pipelined path does not copy it):
 * fetch it with svx_view_read_shaded (host, synchronises) or svx_view_shaded_pointer (device). */
SVX_API int32_t svx_view_set_shading(svx_view* view, const float* light_normal);
SVX_API int32_t svx_view_read_shaded(svx_view* view, uint32_t* rgba8);
SVX_API int32_t svx_view_shaded_pointer(const svx_view* view, void** rgba8);
/* OctreeGPUView::set_resolution / resolution, src/raytracing/bevy/mod.rs:62-88 */
SVX_API int32_t svx_view_set_resolution(svx_view* view, uint32_t width, uint32_t height);
SVX_API int32_t svx_view_resolution(const svx_view* view, uint32_t* width, uint32_t* height);
/* Multi-GPU sharding: this view renders image rows r with (r / rows_per_band) % world == rank. world = 1 renders
 * everything (the default). rows_per_band must be a power of two. */
SVX_API int32_t svx_view_set_shard(svx_view* view, uint32_t rank, uint32_t world, uint32_t rows_per_band);
/* Launch schedule of the viewport kernel: 0 = one CTA per 16x8 pixel block (default), 1 = persistent warps pulling 8x4 tiles
 * from a device-side ticket counter, 2 = persistent warps whose finished lanes take new pixels while the others keep their
 * place in the tree (lane refill; an experiment, 16-48 % slower on every measured scene - profiles/r02_experiments.md section 8).
 * Results are identical. */
SVX_API int32_t svx_view_set_schedule(svx_view* view, int32_t persistent);
/* With compact rows the shard's rows are stored back to back ([rows_local][width], band-major) instead of at their image
 * rows: the layout a collective gather wants. */
SVX_API int32_t svx_view_set_compact_rows(svx_view* view, int32_t enabled);
/* Device pointers of the view's framebuffer planes (valid until set_resolution / free) */
SVX_API int32_t svx_view_frame_pointers(const svx_view* view, void** hit_id, void** albedo, void** distance);
/* ---- Tile-sharded frames over several GPUs, one process per GPU ----------------------------------------------
 * The reference renders on one device (no multi-GPU code); rays are independent and the tree is read-only,
 * so `world` GPUs, each with its own replica of the tree (its own svx_gpu_host), render the rows
 * (row / rows_per_band) % world == rank of ONE frame. The gather is fused into the viewport kernel: a peer's kernel
 * stores its pixels straight into the root view's framebuffer over NVLink (CUDA IPC mapping) and the hand-over runs on
 * device-side flags - no host barrier and no collective per frame. Protocol of one frame (every member calls
 * svx_view_render once, in any order, from its own process):
 *   root : viewport kernel (publishes "go" for this frame, renders rank 0's rows) + a kernel that waits until every
 *          peer's rows have arrived (and, with SVX_WIRE_ID_DISTANCE, resolves their albedo)
 *   peer : a kernel that waits for "go" (so a frame the root's consumer still reads is never overwritten) + viewport
 *          kernel storing into the root's planes; its last CTA publishes "done"
 * After the root's stream has passed the render call the root's framebuffer holds the complete frame, byte-identical to
 * a single-GPU render. A member that does not show up within SVX_GATHER_TIMEOUT_MS (default 5000) makes the waiting
 * side's next synchronising call return SVX_E_TIMEOUT instead of hanging the device.
 *
 * svx_view_gather_open : makes `root` rank 0 of a `world`-way gather and (out != NULL) fills the handle to ship to the
 *                        other processes. Idempotent for the same shape. The root must outlive the membership of peers in OTHER
 *                        processes (they hold CUDA IPC mappings of its frame: close them first).
 * svx_view_gather_join : another process, `rank` in 1..world-1, same resolution as the root.
 * svx_view_gather_join_local : the same for a view of THIS process (any device with peer access to the root's, or the
 *                        root's own device), where CUDA IPC cannot be used.
 * svx_view_gather_close: leaves the gather (root or peer); the view renders whole frames into its own framebuffer again.
 *                        Closing (or freeing) a root also takes its svx_view_gather_join_local peers out of the gather.
 * While a view is a member, set_resolution / set_shard / compact rows / the shaded plane / the pipelined read-back
 * are refused. */
SVX_API int32_t svx_view_gather_open(svx_view* root, uint32_t world, uint32_t rows_per_band, int32_t wire /* svx_wire_format */,
                                     svx_gather_handle* out);
SVX_API int32_t svx_view_gather_join(svx_view* view, uint32_t rank, const svx_gather_handle* handle);
SVX_API int32_t svx_view_gather_join_local(svx_view* view, uint32_t rank, svx_view* root);
SVX_API int32_t svx_view_gather_close(svx_view* view);
/* role: 0 none, 1 root, 2 peer; frames = render calls since the gather was opened / joined. Any output may be NULL. */
SVX_API int32_t svx_view_gather_info(const svx_view* view, int32_t* role, uint32_t* rank, uint32_t* world, uint32_t* frames);
/* One frame: in-kernel ray generation (examples/cpu_render.rs:78-114) + get_by_ray per pixel + framebuffer write.
 * Asynchronous on the view's stream unless `out` is non-null, in which case the call synchronises and fills it. */
SVX_API int32_t svx_view_render(svx_view* view, svx_frame* out);
/* Same frame, delivered into HOST buffers (any may be null): includes the device->host copies. The buffers are
 * full-frame planes [h*w]; a view sharded with svx_view_set_shard copies only the rows it owns into them (strided
 * copies of whole bands), so several GPUs - or processes sharing the host planes - assemble one frame in host memory,
 * each over its own PCIe link. Passing albedo = NULL ships 8 bytes per pixel: albedo is palette[hit_id & 0xFFFF]. */
SVX_API int32_t svx_view_render_to_host(svx_view* view, uint32_t* hit_id, uint32_t* albedo, float* distance);
/* The framebuffer as it stands, without rendering: waits for the view's stream, then copies whole planes (any may be
 * null). On the root of a gather this is the assembled frame of the last svx_view_render. */
SVX_API int32_t svx_view_read_frame(svx_view* view, uint32_t* hit_id, uint32_t* albedo, float* distance);
/* Pipelined variant: returns as soon as the kernel and the copies are queued. Frames alternate between two framebuffer
 * slots; the copies run on a second stream, so frame i's device->host copy overlaps frame i+1's kernel. At most two
 * frames are in flight (a third submission first waits for the oldest). The host buffers must stay valid, and are not
 * complete, until svx_view_wait_host() has retired the frame. Replaces the frame hand-over of
 * OctreeGPUView::output_texture (src/raytracing/bevy/mod.rs:84-88), which the reference leaves to Bevy's render graph. */
SVX_API int32_t svx_view_render_to_host_async(svx_view* view, uint32_t* hit_id, uint32_t* albedo, float* distance);
/* Blocks until at most `keep_in_flight` (0, 1) pipelined frames are outstanding, oldest first. kernel_ms_total
 * (optional) receives and resets the summed CUDA-event kernel time of the frames retired since the last query. */
SVX_API int32_t svx_view_wait_host(svx_view* view, uint32_t keep_in_flight, float* kernel_ms_total);
/* Batch mode: n camera poses through the same view (pose-sharded rendering). Host output arrays are
 * [n][h*w] (any may be null). kernel_ms_total (optional) receives the summed CUDA-event kernel time. */
SVX_API int32_t svx_view_render_batch(svx_view* view, const svx_viewport* poses, uint32_t n, uint32_t* hit_id,
                                      uint32_t* albedo, float* distance, float* kernel_ms_total);
/* Raw handles for interop with an existing CUDA context (torch, NCCL): the view's cudaStream_t and device id */
SVX_API void* svx_view_cuda_stream(const svx_view* view);
SVX_API int32_t svx_view_device(const svx_view* view);
SVX_API int32_t svx_view_synchronize(svx_view* view);
/* Device-side stopwatch on the view's stream: CUDA events recorded around whatever is enqueued between the two calls.
 * svx_view_timer_stop synchronises the stream and returns the elapsed milliseconds. */
SVX_API int32_t svx_view_timer_start(svx_view* view);
SVX_API int32_t svx_view_timer_stop(svx_view* view, float* elapsed_ms);
/* Evicts the L2 cache by overwriting a scratch buffer larger than L2 on the view's stream (benchmark hygiene). */
SVX_API int32_t svx_view_flush_l2(svx_view* view);
/* Number of kernel launches this view/host has issued since creation (bench.py's gpu_launches claim) */
SVX_API uint64_t svx_view_launch_count(const svx_view* view);

/* ---- svx_multi: one process drives several GPUs --------------------------------------------------------- */
/* One replica of the tree (OctreeGPUHost) and one view per listed device; devices[0] assembles the frame. The same
 * device-side protocol as svx_view_gather_* with peer access instead of CUDA IPC. The octree must outlive the handle.
 * (A device may be listed more than once: each entry gets its own replica and stream.) */
SVX_API int32_t svx_multi_create(const svx_octree* tree, const int32_t* devices, uint32_t n_devices, const svx_viewport* viewport,
                                 uint32_t width, uint32_t height, uint32_t rows_per_band, int32_t wire /* svx_wire_format */,
                                 svx_multi** out);
SVX_API void svx_multi_free(svx_multi* multi);
SVX_API uint32_t svx_multi_device_count(const svx_multi* multi);
SVX_API svx_view* svx_multi_view(svx_multi* multi, uint32_t i); /* the gather member on devices[i] (owned by `multi`) */
SVX_API int32_t svx_multi_set_viewport(svx_multi* multi, const svx_viewport* viewport);
SVX_API int32_t svx_multi_set_glass_mode(svx_multi* multi, int32_t mode);
SVX_API int32_t svx_multi_set_viewing_distance(svx_multi* multi, float viewing_distance);
SVX_API int32_t svx_multi_reload(svx_multi* multi); /* svx_gpu_host_reload of every replica after the tree was edited */
/* One frame, tile-sharded over all devices, assembled in devices[0]'s framebuffer. Asynchronous unless `out` is given
 * (then: synchronises devices[0], kernel_ms = the root's viewport kernel plus its wait for the slowest peer). */
SVX_API int32_t svx_multi_render(svx_multi* multi, svx_frame* out);
/* The same frame into HOST planes [h*w] (any may be NULL): every GPU copies the rows it rendered over its own PCIe link,
 * no NVLink hop. Use page-locked buffers (cudaHostRegister with cudaHostRegisterPortable) for overlapping copies. */
SVX_API int32_t svx_multi_render_to_host(svx_multi* multi, uint32_t* hit_id, uint32_t* albedo, float* distance);
/* Batch mode (pose-sharded): pose k is rendered by device k % n into host planes [n_poses][h*w]; nothing is exchanged
 * between the GPUs. kernel_ms_total (optional) = summed kernel time over all devices. */
SVX_API int32_t svx_multi_render_poses(svx_multi* multi, const svx_viewport* poses, uint32_t n_poses, uint32_t* hit_id,
                                       uint32_t* albedo, float* distance, float* kernel_ms_total);

#ifdef __cplusplus
}
#endif
#endif /* SHOCOVOX_B200_H */
