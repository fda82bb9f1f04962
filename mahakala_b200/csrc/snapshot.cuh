// Host-side snapshot handle shared by sample.cu and render.cu.
#pragma once
#include "sample.cuh"

struct mk_snapshot {
    mk::SnapshotView view;
    void* cells;
    double* geom;
    int* grid;
    long cell_bytes;
    long total_bytes;
    int device;
};
