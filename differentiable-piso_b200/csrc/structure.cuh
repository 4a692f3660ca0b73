// structure.cuh -- closed-form CSR layout of the advection-diffusion matrices.
//
// Reproduces, bit-exactly, what calcCsrRowPtrGpu and the slot arithmetic of calcAdvetionMatrixX/Y
// produce (CUDAsrc/central_difference_csr_op.cu.cc:472-505 and :162-231), but is derived independently:
// a row's entries are its existing neighbours plus itself in ascending column order, so the position of
// an entry is simply the number of existing entries with a smaller column, and row_ptr is 5*row minus
// the neighbours missing in the rows before it.
#pragma once
#include "common.cuh"

namespace dpiso {

// component 0 = u (x-staggered, Dx = nx+1, Dy = ny), component 1 = v (y-staggered, Dx = nx, Dy = ny+1)
struct CompDims {
    int Dx, Dy, stag_x, stag_y;
};
DPISO_HD CompDims comp_dims(int ny, int nx, int comp) {
    CompDims c;
    if (comp == 0) { c.Dx = nx + 1; c.Dy = ny; c.stag_x = 1; c.stag_y = 0; }
    else { c.Dx = nx; c.Dy = ny + 1; c.stag_x = 0; c.stag_y = 1; }
    return c;
}

// entries of row (lx, ly): k = 0:x-  1:x+  2:y-  3:y+  4:centre
struct RowLayout {
    int rp;        // row_ptr[row]
    int len;       // entries in the row
    int slot[5];   // position inside the row (valid where has[k]; slot[4] always valid)
    int col[5];    // column index (col[4] = row)
    int has[4];    // structural existence of the neighbour entry
    int reg[4];    // neighbour reached without wrapping (i.e. not across the domain edge)
};

// the periodic wrap skips the duplicated face along the component's own staggered axis
// (":223,230,263,285": dims-1-(d==stag))
DPISO_HD RowLayout row_layout(int lx, int ly, const CompDims &c, int per_x, int per_y) {
    RowLayout r;
    const int Dx = c.Dx, Dy = c.Dy;
    const int row = lx + Dx * ly;
    r.reg[0] = lx > 0; r.reg[1] = lx < Dx - 1; r.reg[2] = ly > 0; r.reg[3] = ly < Dy - 1;
    r.has[0] = r.reg[0] || per_x; r.has[1] = r.reg[1] || per_x;
    r.has[2] = r.reg[2] || per_y; r.has[3] = r.reg[3] || per_y;
    r.col[0] = r.reg[0] ? row - 1 : row + (Dx - 1 - c.stag_x);
    r.col[1] = r.reg[1] ? row + 1 : row - (Dx - 1 - c.stag_x);
    r.col[2] = r.reg[2] ? row - Dx : row + Dx * (Dy - 1 - c.stag_y);
    r.col[3] = r.reg[3] ? row + Dx : row - Dx * (Dy - 1 - c.stag_y);
    r.col[4] = row;
    r.len = 1 + r.has[0] + r.has[1] + r.has[2] + r.has[3];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        int s = 0;
#pragma unroll
        for (int j = 0; j < 5; j++) {
            const int exists = j == 4 ? 1 : r.has[j];
            s += (j != k && exists && r.col[j] < r.col[k]) ? 1 : 0;
        }
        r.slot[k] = s;
    }
    // neighbours missing in the rows strictly before this one
    const int miss_x = 2 * ly + (lx > 0 ? 1 : 0);
    const int miss_y = (ly == 0 ? lx : Dx) + (ly == Dy - 1 ? lx : 0);
    r.rp = 5 * row - (1 - per_x) * miss_x - (1 - per_y) * miss_y;
    return r;
}

}  // namespace dpiso
