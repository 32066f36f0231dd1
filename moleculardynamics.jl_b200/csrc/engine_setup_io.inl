// engine_setup_io.inl -- host side of the steps either side of the hot path: trajectory frames
// (device packing, copy stream, background LAMMPS writer), initialize_velocities / initialize_random on the device, the
// exact binary checkpoint, and the FP64 throughput probe.  Included at the end of engine.cu (same translation unit: it
// uses the Engine struct and the CU / fail helpers defined there); kernels are in setup_io.cuh.
// ------------------------------------------------------------------------------------------------
// Trajectory frames.  Replaces the positions/images reads of write_to_file_lammps
// (src/io.jl:78-170) at its two call sites in the step loop (src/simulation.jl:139-171): the frame is packed on the
// device in the caller's particle order with the unwrapped coordinates already formed, copied to pinned host memory on
// a second stream, and formatted by a background thread of the library, so the step loop goes on while a frame
// is in flight.  Slots are double-buffered (MDB_FRAME_SLOTS); a slot is busy from capture until its file is written
// (or, without a write request, until the next capture into it).
// ------------------------------------------------------------------------------------------------
struct FrameJob {
    int slot;
    std::string path;
    int64_t step;
    int append;
};
struct FrameIO {
    double *d_frame[MDB_FRAME_SLOTS] = {};
    double *h_frame[MDB_FRAME_SLOTS] = {};
    // x-slab handles: rows are the rank's owned particles in slot order, ids travel beside them, one file per rank
    int32_t *d_ids[MDB_FRAME_SLOTS] = {};
    int32_t *h_ids[MDB_FRAME_SLOTS] = {};
    int64_t count[MDB_FRAME_SLOTS] = {};
    bool slab = false;
    int rank = 0;
    cudaEvent_t copied[MDB_FRAME_SLOTS] = {};
    bool captured[MDB_FRAME_SLOTS] = {};
    int writing[MDB_FRAME_SLOTS] = {};  // jobs queued or running on the slot
    cudaEvent_t packed = nullptr;
    cudaStream_t copy_stream = nullptr;
    int64_t n = 0;
    int width = 0;
    std::thread worker;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<FrameJob> jobs;
    bool stop = false;
    std::string io_error;
    // what the writer needs of the engine
    int device = 0, dim = 3;
    double U[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // cell matrix, lattice vectors in the columns
};

// write_to_file_lammps for a diagonal cell: same header lines, same "%lf" columns (src/io.jl:97-167)
static bool write_lammps_frame(const FrameIO *f, const FrameJob &job, std::string &err)
{
    // slab handles write "<path>.<rank>": a complete LAMMPS dump of the rank's own atoms (the "%"-per-processor multi-file
    // convention of LAMMPS' dump command; readers glob <path>.*)
    const std::string path = f->slab ? job.path + "." + std::to_string(f->rank) : job.path;
    FILE *fp = fopen(path.c_str(), job.append ? "a" : "w");
    if (!fp) {
        err = "cannot open " + path;
        return false;
    }
    const int dim = f->dim, W = f->width;
    const int64_t n = f->slab ? f->count[job.slot] : f->n;
    const int32_t *ids = f->slab ? f->h_ids[job.slot] : nullptr;
    fprintf(fp, "ITEM: TIMESTEP\n%lld\n", (long long)job.step);
    fprintf(fp, "ITEM: NUMBER OF ATOMS\n%lld\n", (long long)n);
    // box bounds = norms of the lattice vectors (columns), tilt factors xy = U[1,2], xz = U[1,3], yz = U[2,3] (src/io.jl:104-128)
    const double *U = f->U;
    auto coln = [&](int c) { return std::sqrt(U[c] * U[c] + U[3 + c] * U[3 + c] + U[6 + c] * U[6 + c]); };
    if (dim == 2) {
        fprintf(fp, "ITEM: BOX BOUNDS xy pp pp\n");
        fprintf(fp, "%lf %lf %lf\n", 0.0, std::sqrt(U[0] * U[0] + U[3] * U[3]), U[1]);
        fprintf(fp, "%lf %lf 0.0\n", 0.0, std::sqrt(U[1] * U[1] + U[4] * U[4]));
        fprintf(fp, "%lf %lf 0.0\n", 0.0, 1.0);
        fprintf(fp, "ITEM: ATOMS id type radius x y xu yu\n");
    } else {
        fprintf(fp, "ITEM: BOX BOUNDS xy xz yz pp pp pp\n");
        fprintf(fp, "%lf %lf %lf\n", 0.0, coln(0), U[1]);
        fprintf(fp, "%lf %lf %lf\n", 0.0, coln(1), U[5]);
        fprintf(fp, "%lf %lf %lf\n", 0.0, coln(2), U[2]);
        fprintf(fp, "ITEM: ATOMS id type radius x y z xu yu zu\n");
    }
    // rows are formatted in parallel blocks and written in order
    const double *fr = f->h_frame[job.slot];
    const int nthr = (int)std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
    const int64_t block = 1 << 18;
    std::vector<std::string> bufs(nthr);
    bool ok = true;
    for (int64_t b0 = 0; b0 < n && ok; b0 += block * nthr) {
        std::vector<std::thread> th;
        for (int t = 0; t < nthr; t++) {
            const int64_t lo = b0 + t * block, hi = std::min(n, lo + block);
            bufs[t].clear();
            if (lo >= hi) continue;
            th.emplace_back([&, t, lo, hi]() {
                std::string &out = bufs[t];
                out.reserve((size_t)(hi - lo) * (16 + 14 * (size_t)W));
                // "%lf" of a finite double prints at most 1 + 309 + 1 + 6 characters (a coordinate near 1e308 after a blow-up
                // that has not tripped MDB_ERR_NONFINITE yet): every column is formatted on its own and appended, so no
                // row length can run past the buffer
                char col[352];
                for (int64_t i = lo; i < hi; i++) {
                    const double *r = fr + i * W;
                    int len = snprintf(col, sizeof(col), "%lld %d", (long long)((ids ? (int64_t)ids[i] : i) + 1), 1);
                    out.append(col, (size_t)std::min<int>(len, (int)sizeof(col) - 1));
                    for (int c = 0; c < W; c++) {
                        len = snprintf(col, sizeof(col), " %lf", r[c]);
                        out.append(col, (size_t)std::min<int>(std::max(len, 0), (int)sizeof(col) - 1));
                    }
                    out.push_back('\n');
                }
            });
        }
        for (auto &t : th) t.join();
        for (int t = 0; t < nthr && ok; t++)
            if (!bufs[t].empty() && fwrite(bufs[t].data(), 1, bufs[t].size(), fp) != bufs[t].size()) ok = false;
    }
    if (fclose(fp) != 0) ok = false;
    if (!ok) err = "short write to " + path;
    return ok;
}

static void frame_worker(FrameIO *f)
{
    cudaSetDevice(f->device);
    for (;;) {
        FrameJob job;
        {
            std::unique_lock<std::mutex> lk(f->mu);
            f->cv.wait(lk, [&] { return f->stop || !f->jobs.empty(); });
            if (f->jobs.empty()) return;  // stop requested and nothing left
            job = f->jobs.front();
            f->jobs.pop_front();
        }
        std::string err;
        cudaError_t ce = cudaEventSynchronize(f->copied[job.slot]);
        if (ce != cudaSuccess) err = std::string("frame copy: ") + cudaGetErrorString(ce);
        else write_lammps_frame(f, job, err);
        {
            std::lock_guard<std::mutex> lk(f->mu);
            if (!err.empty() && f->io_error.empty()) f->io_error = err;
            f->writing[job.slot]--;
        }
        f->cv.notify_all();
    }
}

static void free_frames(Engine *e)
{
    FrameIO *f = e->fio;
    if (!f) return;
    {
        std::lock_guard<std::mutex> lk(f->mu);
        f->stop = true;
    }
    f->cv.notify_all();
    if (f->worker.joinable()) f->worker.join();
    if (f->copy_stream) cudaStreamSynchronize(f->copy_stream);
    for (int q = 0; q < MDB_FRAME_SLOTS; q++) {
        cudaFree(f->d_frame[q]);
        if (f->h_frame[q]) cudaFreeHost(f->h_frame[q]);
        cudaFree(f->d_ids[q]);
        if (f->h_ids[q]) cudaFreeHost(f->h_ids[q]);
        if (f->copied[q]) cudaEventDestroy(f->copied[q]);
    }
    if (f->packed) cudaEventDestroy(f->packed);
    if (f->copy_stream) cudaStreamDestroy(f->copy_stream);
    delete f;
    e->fio = nullptr;
}

static int ensure_frames(Engine *e)
{
    const int64_t rows = e->slab ? (int64_t)e->cap_own : e->N;
    if (e->fio && e->fio->n == rows && e->fio->dim == e->dim) return MDB_OK;
    free_frames(e);
    FrameIO *f = new FrameIO();
    e->fio = f;
    f->n = rows;
    f->slab = e->slab;
    f->rank = e->rank;
    f->dim = e->dim;
    f->width = 2 * e->dim + 1;
    f->device = e->cfg.device;
    memcpy(f->U, e->U, sizeof(f->U));
    if (e->dim == 2) f->U[8] = 0.0;  // the writer's 3x3 boxmat of a 2-D cell has an empty third column (src/io.jl:101-102)
    const size_t bytes = sizeof(double) * (size_t)f->n * f->width;
    for (int q = 0; q < MDB_FRAME_SLOTS; q++) {
        CU(cudaMalloc(&f->d_frame[q], bytes));
        CU(cudaMallocHost(&f->h_frame[q], bytes));
        if (e->slab) {
            CU(cudaMalloc(&f->d_ids[q], sizeof(int32_t) * (size_t)rows));
            CU(cudaMallocHost(&f->h_ids[q], sizeof(int32_t) * (size_t)rows));
        }
        CU(cudaEventCreateWithFlags(&f->copied[q], cudaEventDisableTiming));
    }
    CU(cudaEventCreateWithFlags(&f->packed, cudaEventDisableTiming));
    CU(cudaStreamCreateWithFlags(&f->copy_stream, cudaStreamNonBlocking));
    f->worker = std::thread(frame_worker, f);
    return MDB_OK;
}

MDB_EXPORT int mdb_frame_capture(mdb_handle e, int32_t slot)
{
    if (!e) return MDB_ERR_INVALID_ARG;
    if (slot < 0 || slot >= MDB_FRAME_SLOTS) return fail(e, MDB_ERR_INVALID_ARG, "frame slot out of range");
    if (!e->uploaded) return fail(e, MDB_ERR_STATE, "nothing uploaded");
    CU(cudaSetDevice(e->cfg.device));
    int rc = ensure_frames(e);
    if (rc) return rc;
    FrameIO *f = e->fio;
    if (e->slab) {  // the owned count changes at rebuilds: read it (rank-local call, no collective)
        if ((rc = sync_ctl(e))) return rc;
        e->n = e->h_ctl->n_own;
        e->stats.n_owned = e->n;
    }
    {
        // the slot's host image must be on disk before it is overwritten
        std::unique_lock<std::mutex> lk(f->mu);
        f->cv.wait(lk, [&] { return f->writing[slot] == 0; });
    }
    cudaStream_t s = e->stream;
    // (the previous copy out of this slot's device buffer was awaited above or by mdb_frame_wait; order it anyway)
    if (f->captured[slot]) CU(cudaStreamWaitEvent(s, f->copied[slot], 0));
    const int blocks = std::max(1, std::min(nblk(e->n, kStreamBlock), e->nsm * 8));
    if (e->slab) {
        if (e->dim == 3) k_pack_frame_slab<3><<<blocks, kStreamBlock, 0, s>>>(e->ctl, e->grid, f->d_frame[slot], f->d_ids[slot]);
        else k_pack_frame_slab<2><<<blocks, kStreamBlock, 0, s>>>(e->ctl, e->grid, f->d_frame[slot], f->d_ids[slot]);
        f->count[slot] = e->n;
    } else if (e->dim == 3) k_pack_frame<3><<<blocks, kStreamBlock, 0, s>>>(e->n, e->ctl, e->grid, f->d_frame[slot]);
    else k_pack_frame<2><<<blocks, kStreamBlock, 0, s>>>(e->n, e->ctl, e->grid, f->d_frame[slot]);
    e->stats.kernel_launches += 1;
    CU(cudaEventRecord(f->packed, s));
    CU(cudaStreamWaitEvent(f->copy_stream, f->packed, 0));
    const int64_t rows = e->slab ? (int64_t)e->n : f->n;
    CU(cudaMemcpyAsync(f->h_frame[slot], f->d_frame[slot], sizeof(double) * (size_t)rows * f->width, cudaMemcpyDeviceToHost, f->copy_stream));
    if (e->slab) CU(cudaMemcpyAsync(f->h_ids[slot], f->d_ids[slot], sizeof(int32_t) * (size_t)rows, cudaMemcpyDeviceToHost, f->copy_stream));
    CU(cudaEventRecord(f->copied[slot], f->copy_stream));
    f->captured[slot] = true;
    CU(cudaGetLastError());
    return MDB_OK;
}

MDB_EXPORT int mdb_frame_wait(mdb_handle e, int32_t slot, const double **frame, int32_t *width)
{
    if (!e) return MDB_ERR_INVALID_ARG;
    if (slot < 0 || slot >= MDB_FRAME_SLOTS || !e->fio || !e->fio->captured[slot]) return fail(e, MDB_ERR_STATE, "no frame captured in this slot");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaEventSynchronize(e->fio->copied[slot]));
    if (frame) *frame = e->fio->h_frame[slot];
    if (width) *width = e->fio->width;
    return MDB_OK;
}

MDB_EXPORT int mdb_frame_write_lammps(mdb_handle e, int32_t slot, const char *path, int64_t step, int32_t append)
{
    if (!e || !path) return MDB_ERR_INVALID_ARG;
    if (slot < 0 || slot >= MDB_FRAME_SLOTS || !e->fio || !e->fio->captured[slot]) return fail(e, MDB_ERR_STATE, "no frame captured in this slot");
    FrameIO *f = e->fio;
    {
        std::lock_guard<std::mutex> lk(f->mu);
        f->jobs.push_back(FrameJob{slot, std::string(path), step, append});
        f->writing[slot]++;
    }
    f->cv.notify_all();
    return MDB_OK;
}

MDB_EXPORT int mdb_frame_flush(mdb_handle e)
{
    if (!e) return MDB_ERR_INVALID_ARG;
    FrameIO *f = e->fio;
    if (!f) return MDB_OK;
    std::unique_lock<std::mutex> lk(f->mu);
    f->cv.wait(lk, [&] {
        for (int q = 0; q < MDB_FRAME_SLOTS; q++)
            if (f->writing[q]) return false;
        return true;
    });
    if (!f->io_error.empty()) {
        std::string msg = f->io_error;
        f->io_error.clear();
        lk.unlock();
        return fail(e, MDB_ERR_IO, msg);
    }
    return MDB_OK;
}

// ------------------------------------------------------------------------------------------------
// initialize_velocities on the device (src/initialization.jl:32-47)
// ------------------------------------------------------------------------------------------------
MDB_EXPORT int mdb_init_velocities(mdb_handle e, double ktemp, uint64_t stream)
{
    if (!e) return MDB_ERR_INVALID_ARG;
    if (!e->uploaded) return fail(e, MDB_ERR_STATE, "mdb_upload first");
    if (!(ktemp > 0) || e->N < 2) return fail(e, MDB_ERR_INVALID_ARG, "ktemp must be > 0 and n_particles >= 2");
    CU(cudaSetDevice(e->cfg.device));
    // x-slabs: collective over the ring (every rank, or the rank-0 handle of an in-process ring).  The normals are keyed by
    // particle id, so every rank draws exactly what a single domain would draw for its particles; the centre-of-mass and
    // temperature sums are all-reduced between the sweeps.
    Group storage, *G = nullptr;
    if (e->slab) {
        int rc = slab_group(e, storage, &G);
        if (rc) return rc;
        for (Engine *m : *G)
            if (!m->uploaded) return fail(e, MDB_ERR_STATE, "mdb_upload has not been called on every slab");
    } else {
        storage.assign(1, e);
        G = &storage;
    }
    for (int stage = 0; stage < 3; stage++) {
        for (Engine *m : *G) {
            const int64_t n = m->slab ? -1 : (int64_t)m->n;
            const int blocks = std::max(1, std::min(std::min(nblk(grid_particles(m), kStreamBlock), m->nsm * 8), kMaxPartials));
            if (m->dim == 3) k_vel_init<3><<<blocks, kStreamBlock, 0, m->stream>>>(stage, n, m->cfg.seed, stream, m->ctl, m->part);
            else k_vel_init<2><<<blocks, kStreamBlock, 0, m->stream>>>(stage, n, m->cfg.seed, stream, m->ctl, m->part);
            if (stage < 2) k_vel_reduce<<<1, kStreamBlock, 0, m->stream>>>(stage, blocks, m->part, (double)m->N, m->dim, ktemp, m->ctl, m->slab ? 1 : 0);
            m->stats.kernel_launches += stage < 2 ? 2 : 1;
        }
        if (e->slab && stage < 2) {
            int rc = group_allreduce(*G, 4, false, [](Engine *m) { return m->ctl->red; });
            if (rc) return rc;
            for (Engine *m : *G) {
                k_vel_reduce<<<1, kStreamBlock, 0, m->stream>>>(stage, 0, m->part, (double)m->N, m->dim, ktemp, m->ctl, 2);
                m->stats.kernel_launches += 1;
            }
        }
    }
    for (Engine *m : *G) {
        CU(cudaStreamSynchronize(m->stream));
        m->have_vel = true;
    }
    CU(cudaGetLastError());
    return MDB_OK;
}

// initialize_random's first half (src/initialization.jl:20-27): uniform random positions in the cell, drawn on the
// device.  The second half (Packmol's overlap removal) is mdb_fire_minimize on a handle created with MDB_POT_SOFT.
MDB_EXPORT int mdb_random_positions(mdb_handle e, uint64_t stream)
{
    if (!e) return MDB_ERR_INVALID_ARG;
    if (!e->uploaded) return fail(e, MDB_ERR_STATE, "mdb_upload first (diameters; the positions passed there are replaced)");
    if (e->slab) return fail(e, MDB_ERR_STATE, "nranks > 1: draw the positions on one handle and upload them");
    CU(cudaSetDevice(e->cfg.device));
    cudaStream_t s = e->stream;
    const int blocks = std::max(1, std::min(nblk(e->n, kStreamBlock), e->nsm * 8));
    if (e->dim == 3) k_random_positions<3><<<blocks, kStreamBlock, 0, s>>>(e->n, e->grid, e->cfg.seed, stream, e->ctl);
    else k_random_positions<2><<<blocks, kStreamBlock, 0, s>>>(e->n, e->grid, e->cfg.seed, stream, e->ctl);
    e->stats.kernel_launches += 1;
    CU(cudaStreamSynchronize(s));
    CU(cudaGetLastError());
    return MDB_OK;
}

// ------------------------------------------------------------------------------------------------
// Exact binary checkpoint.  The file holds the state in DEVICE SLOT ORDER together with the RNG
// step counter; saving also invalidates the resident Verlet list, so the next step of the run that was saved and the
// first step of a run restored from the file start from the same rebuild of the same slot order: the continuation is
// bit-identical (tests/test_gpu_setup_io.py).
// ------------------------------------------------------------------------------------------------
struct CkptHeader {
    char magic[8];
    uint32_t version;
    int32_t dim;
    int64_t n_particles;
    double unitcell[9];
    uint64_t seed;
    uint64_t rng_step;
    int32_t have_vel;
    int32_t reserved[7];
};
static const char kCkptMagic[8] = {'M', 'D', 'B', '2', '0', '0', 'C', 'K'};
// x-slab handles write one file per rank, "<path>.<rank>", version 2: the header above followed by this block.  The global
// diameter range travels with every file so that a fresh handle can plan its slab (cell grid, capacities, mailbox) from
// the file alone; restoring is collective (every rank loads its file, then the ring reconnects at the next run).
struct CkptSlab {
    int32_t rank, nranks;
    int64_t n_local;
    double smin, smax;
    int64_t reserved[4];
};

struct CkptBuffers {
    double4 *pos = nullptr;
    double *vel = nullptr, *frc = nullptr;
    int32_t *img = nullptr, *id = nullptr;
    ~CkptBuffers()
    {
        cudaFree(pos); cudaFree(vel); cudaFree(frc); cudaFree(img); cudaFree(id);
    }
    cudaError_t alloc(int64_t n, int d)
    {
        cudaError_t ce;
        if ((ce = cudaMalloc(&pos, sizeof(double4) * n)) != cudaSuccess) return ce;
        if ((ce = cudaMalloc(&vel, sizeof(double) * n * d)) != cudaSuccess) return ce;
        if ((ce = cudaMalloc(&frc, sizeof(double) * n * d)) != cudaSuccess) return ce;
        if ((ce = cudaMalloc(&img, sizeof(int32_t) * n * d)) != cudaSuccess) return ce;
        return cudaMalloc(&id, sizeof(int32_t) * n);
    }
};

MDB_EXPORT int mdb_checkpoint_save(mdb_handle e, const char *path)
{
    if (!e || !path) return MDB_ERR_INVALID_ARG;
    if (!e->uploaded) return fail(e, MDB_ERR_STATE, "nothing uploaded");
    CU(cudaSetDevice(e->cfg.device));
    cudaStream_t s = e->stream;
    if (e->slab) {  // rank-local: the owned count changes at rebuilds
        int rc0 = sync_ctl(e);
        if (rc0) return rc0;
        e->n = e->h_ctl->n_own;
        e->stats.n_owned = e->n;
    }
    const std::string file = e->slab ? std::string(path) + "." + std::to_string(e->rank) : std::string(path);
    const int64_t n = e->n;
    const int d = e->dim;
    CkptBuffers b;
    CU(b.alloc(n, d));
    const int blocks = std::max(1, std::min(nblk(n, kStreamBlock), e->nsm * 8));
    if (d == 3) k_ckpt_pack<3><<<blocks, kStreamBlock, 0, s>>>(n, e->ctl, b.pos, b.vel, b.frc, b.img, b.id);
    else k_ckpt_pack<2><<<blocks, kStreamBlock, 0, s>>>(n, e->ctl, b.pos, b.vel, b.frc, b.img, b.id);
    e->stats.kernel_launches += 1;
    std::vector<double4> hpos(n);
    std::vector<double> hvel((size_t)n * d), hfrc((size_t)n * d);
    std::vector<int32_t> himg((size_t)n * d), hid(n);
    CU(cudaMemcpyAsync(hpos.data(), b.pos, sizeof(double4) * n, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(hvel.data(), b.vel, sizeof(double) * n * d, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(hfrc.data(), b.frc, sizeof(double) * n * d, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(himg.data(), b.img, sizeof(int32_t) * n * d, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(hid.data(), b.id, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
    // both futures (this engine going on, an engine restored from the file) rebuild the list at their next step
    CU(cudaMemsetAsync(&e->ctl->list_valid, 0, sizeof(int), s));
    int rc = sync_ctl(e);
    if (rc) return rc;
    CkptHeader h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, kCkptMagic, 8);
    h.version = e->slab ? 2 : 1;
    h.dim = d;
    h.n_particles = e->N;
    memcpy(h.unitcell, e->cfg.unitcell, sizeof(h.unitcell));
    h.seed = e->cfg.seed;
    h.rng_step = e->h_ctl->rng_step;
    h.have_vel = e->have_vel ? 1 : 0;
    FILE *fp = fopen(file.c_str(), "wb");
    if (!fp) return fail(e, MDB_ERR_IO, std::string("cannot open ") + file);
    bool ok = fwrite(&h, sizeof(h), 1, fp) == 1;
    if (e->slab) {
        CkptSlab sl;
        memset(&sl, 0, sizeof(sl));
        sl.rank = e->rank; sl.nranks = e->nranks; sl.n_local = n; sl.smin = e->smin; sl.smax = e->smax;
        ok = ok && fwrite(&sl, sizeof(sl), 1, fp) == 1;
    }
    ok = ok && fwrite(hpos.data(), sizeof(double4), n, fp) == (size_t)n &&
              fwrite(hvel.data(), sizeof(double), (size_t)n * d, fp) == (size_t)n * d &&
              fwrite(hfrc.data(), sizeof(double), (size_t)n * d, fp) == (size_t)n * d &&
              fwrite(himg.data(), sizeof(int32_t), (size_t)n * d, fp) == (size_t)n * d &&
              fwrite(hid.data(), sizeof(int32_t), n, fp) == (size_t)n;
    ok = (fclose(fp) == 0) && ok;
    if (!ok) return fail(e, MDB_ERR_IO, std::string("short write to ") + file);
    return MDB_OK;
}

MDB_EXPORT int mdb_checkpoint_load(mdb_handle e, const char *path)
{
    if (!e || !path) return MDB_ERR_INVALID_ARG;
    CU(cudaSetDevice(e->cfg.device));
    // x-slab handles read "<path>.<rank>" (collective: every rank restores its own file; the ring reconnects at the next run)
    const std::string file = e->slab ? std::string(path) + "." + std::to_string(e->rank) : std::string(path);
    FILE *fp = fopen(file.c_str(), "rb");
    if (!fp) return fail(e, MDB_ERR_IO, std::string("cannot open ") + file);
    CkptHeader h;
    if (fread(&h, sizeof(h), 1, fp) != 1 || memcmp(h.magic, kCkptMagic, 8) != 0 || h.version != (uint32_t)(e->slab ? 2 : 1)) {
        fclose(fp);
        return fail(e, MDB_ERR_IO, file + (e->slab ? " is not an mdb200 slab checkpoint" : " is not an mdb200 checkpoint"));
    }
    if (h.dim != e->dim || h.n_particles != e->N || memcmp(h.unitcell, e->cfg.unitcell, sizeof(h.unitcell)) != 0) {
        fclose(fp);
        return fail(e, MDB_ERR_INVALID_ARG, "checkpoint was written for a different system (dimension, particle count or unit cell)");
    }
    CkptSlab sl;
    memset(&sl, 0, sizeof(sl));
    if (e->slab) {
        if (fread(&sl, sizeof(sl), 1, fp) != 1 || sl.rank != e->rank || sl.nranks != e->nranks || sl.n_local < 0 || sl.n_local > e->N) {
            fclose(fp);
            return fail(e, MDB_ERR_INVALID_ARG, file + " was written by another rank or for another number of slabs");
        }
    }
    const int64_t n = e->slab ? sl.n_local : e->N;
    const int d = e->dim;
    std::vector<double4> hpos(n);
    std::vector<double> hvel((size_t)n * d), hfrc((size_t)n * d);
    std::vector<int32_t> himg((size_t)n * d), hid(n);
    bool ok = fread(hpos.data(), sizeof(double4), n, fp) == (size_t)n && fread(hvel.data(), sizeof(double), (size_t)n * d, fp) == (size_t)n * d &&
              fread(hfrc.data(), sizeof(double), (size_t)n * d, fp) == (size_t)n * d &&
              fread(himg.data(), sizeof(int32_t), (size_t)n * d, fp) == (size_t)n * d && fread(hid.data(), sizeof(int32_t), n, fp) == (size_t)n;
    fclose(fp);
    if (!ok) return fail(e, MDB_ERR_IO, std::string("truncated checkpoint ") + file);
    double smin = e->slab ? sl.smin : 0.0, smax = e->slab ? sl.smax : 0.0;
    if (!e->slab) {
        smin = smax = hpos[0].w;
        for (int64_t i = 1; i < n; i++) {
            smin = std::min(smin, hpos[i].w);
            smax = std::max(smax, hpos[i].w);
        }
    }
    if (!(smin > 0) || !std::isfinite(smax)) return fail(e, MDB_ERR_IO, "checkpoint holds invalid diameters");
    // the kernels scatter through id[] (export, frames, velocity import): it must be a permutation of 0..n-1 (a subset
    // without repeats for a slab), and a state with non-finite entries is a corrupted file, not something to step
    {
        std::vector<uint8_t> seen((size_t)e->N, 0);
        for (int64_t i = 0; i < n; i++) {
            const int32_t q = hid[i];
            if (q < 0 || q >= e->N || seen[q])
                return fail(e, MDB_ERR_IO, std::string("corrupted checkpoint ") + file + ": particle ids are not a permutation of 0..n-1");
            seen[q] = 1;
        }
        bool finite = true;
        for (int64_t i = 0; i < n && finite; i++) finite = std::isfinite(hpos[i].x) && std::isfinite(hpos[i].y) && std::isfinite(hpos[i].z);
        for (size_t i = 0; i < (size_t)n * d && finite; i++) finite = std::isfinite(hvel[i]) && std::isfinite(hfrc[i]);
        if (!finite) return fail(e, MDB_ERR_IO, std::string("corrupted checkpoint ") + file + ": non-finite positions, velocities or forces");
    }
    e->smin = smin; e->smax = smax;
    int rc;
    if ((rc = plan_neighbors(e))) return rc;
    e->n = (int)n;
    int64_t need_cap = n;
    if (e->slab) {  // the capacities mdb_upload would have planned
        const int64_t n_est = std::max<int64_t>(n, e->N / e->nranks);
        e->cap_own = (int)(((int64_t)(n_est * 1.25) + 4096 + 31) & ~(int64_t)31);
        need_cap = e->cap_own;
    }
    if (e->cap < need_cap || !e->st[0].pos) {
        if ((rc = alloc_state(e, need_cap))) return rc;
    }
    if (e->slab) e->cap_own = (int)e->cap;
    if ((rc = ensure_stage(e, std::max<int64_t>(n, 1)))) return rc;
    if ((rc = alloc_neighbors(e))) return rc;
    drop_graph(e);
    if (e->slab && (rc = alloc_slab(e))) return rc;
    if (e->group && (*e->group)[0]) drop_graph((*e->group)[0]);
    if (d == 3) query_occupancy<3>(e);
    else query_occupancy<2>(e);
    cudaStream_t s = e->stream;
    CkptBuffers b;
    CU(b.alloc(std::max<int64_t>(n, 1), d));
    CU(cudaMemcpyAsync(b.pos, hpos.data(), sizeof(double4) * n, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(b.vel, hvel.data(), sizeof(double) * n * d, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(b.frc, hfrc.data(), sizeof(double) * n * d, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(b.img, himg.data(), sizeof(int32_t) * n * d, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(b.id, hid.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
    e->rng_step = h.rng_step;
    DevCtl c;
    memset(&c, 0, sizeof(c));
    c.alpha = 1.0;
    c.rng_step = e->rng_step;
    c.n_own = (int)n;
    c.n_tmp = (int)n;
    c.st[0] = e->st[0];
    c.st[1] = e->st[1];
    c.epoch = 0;
    c.gpos_m = e->grid.gpos_m;
    *e->h_ctl = c;
    CU(cudaMemcpyAsync(e->ctl, e->h_ctl, sizeof(DevCtl), cudaMemcpyHostToDevice, s));
    const int blocks = std::max(1, std::min(nblk(n, kStreamBlock), e->nsm * 8));
    if (d == 3) k_ckpt_unpack<3><<<blocks, kStreamBlock, 0, s>>>(n, e->st[0], b.pos, b.vel, b.frc, b.img, b.id);
    else k_ckpt_unpack<2><<<blocks, kStreamBlock, 0, s>>>(n, e->st[0], b.pos, b.vel, b.frc, b.img, b.id);
    e->stats.kernel_launches += 1;
    CU(cudaStreamSynchronize(s));
    CU(cudaGetLastError());
    e->uploaded = true;
    e->have_vel = h.have_vel != 0;
    e->stats.n_owned = e->n;
    return MDB_OK;
}

MDB_EXPORT int mdb_force_kernel_info(mdb_handle e, int32_t info[6])
{
    if (!e || !info) return MDB_ERR_INVALID_ARG;
    CU(cudaSetDevice(e->cfg.device));
    cudaFuncAttributes a;
    memset(&a, 0, sizeof(a));
    cudaError_t ce = cudaErrorInvalidValue;
    int var = 0;
    dispatch_pot(e->cfg.potential, [&](auto pot) {
        typedef decltype(pot) Pot;
        var = (e->tri || !has_force_variants<Pot>()) ? 0 : e->force_variant;
        if constexpr (has_force_variants<Pot>()) {
            if (var == 1) {
                ce = e->dim == 3 ? cudaFuncGetAttributes(&a, k_force_list_staged<3, Pot, 2, false>) : cudaFuncGetAttributes(&a, k_force_list_staged<2, Pot, 2, false>);
                return;
            }
            if (var == 2) {
                ce = e->dim == 3 ? cudaFuncGetAttributes(&a, k_force_list_tma<3, Pot, 2, false>) : cudaFuncGetAttributes(&a, k_force_list_tma<2, Pot, 2, false>);
                return;
            }
        }
        var = 0;
        ce = e->dim == 3 ? cudaFuncGetAttributes(&a, k_force_list<3, Pot, 2, false, false>) : cudaFuncGetAttributes(&a, k_force_list<2, Pot, 2, false, false>);
    });
    if (ce != cudaSuccess) return fail(e, MDB_ERR_CUDA, std::string("cudaFuncGetAttributes: ") + cudaGetErrorString(ce));
    info[0] = a.numRegs;
    info[1] = (int32_t)a.sharedSizeBytes;
    info[2] = e->force_cta_per_sm;
    info[3] = var;
    info[4] = kForceBlock;
    info[5] = (int32_t)a.localSizeBytes;
    return MDB_OK;
}

// measured FP64 (DFMA) throughput of this device in TFLOP/s: denominator for FP64-pipe utilisation
MDB_EXPORT int mdb_measure_fp64_peak(mdb_handle e, double *tflops)
{
    if (!e || !tflops) return MDB_ERR_INVALID_ARG;
    CU(cudaSetDevice(e->cfg.device));
    cudaStream_t s = e->stream;
    const int iters = 1 << 14, blocks = e->nsm * 8, threads = 256;
    k_fp64_probe<<<blocks, threads, 0, s>>>(256, 1.0, e->d_scratch);  // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        CU(cudaEventRecord(e->evf0, s));
        k_fp64_probe<<<blocks, threads, 0, s>>>(iters, 1.0, e->d_scratch);
        CU(cudaEventRecord(e->evf1, s));
        CU(cudaEventSynchronize(e->evf1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e->evf0, e->evf1));
        const double flops = 2.0 * 8.0 * (double)iters * (double)blocks * threads;
        best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    e->stats.kernel_launches += 6;
    CU(cudaGetLastError());
    *tflops = best;
    return MDB_OK;
}
