/* A plain C consumer of include/fvb200.h (no C++, no Python, no torch): proves the boundary is a C ABI.
 * Built and run by tests/test_layout.py with gcc; calls only host-side entry points (no GPU needed). */
#include <stdio.h>
#include <string.h>

#include "fvb200.h"

int main(void) {
  fvb_yolo_geom g;
  int l;
  memset(&g, 0, sizeof g);
  g.levels = 3;
  g.batch = 8;
  g.anchors = 3;
  g.channels = 85;
  for (l = 0; l < 3; ++l) {
    g.height[l] = g.width[l] = 13 << l;
    g.stride[l] = (float)(32 >> l);
  }
  g.head_layout = FVB_HEAD_BAHWK;
  if (fvb_abi_version() != FVB_ABI_VERSION) return 1;
  if (fvb_yolo_rows_per_image(&g) != 10647) return 2;
  if (fvb_yolo_bitmap_words(&g) != 333) return 3;
  g.channels = 3; /* invalid: the error comes back as a code + message, not an exception */
  if (fvb_yolo_rows_per_image(&g) >= 0) return 4;
  if (strstr(fvb_last_error(), "channels") == NULL) return 5;
  printf("abi ok, %zu-byte geom\n", sizeof g);
  return 0;
}
