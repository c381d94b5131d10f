"""include/alm_ocr.h is the contract: it must compile as C99 and as C++17 on its own, and a freestanding C host must
link against libalm_ocr.so and reach the host-only entry points (no GPU involved)."""
import os
import shutil
import subprocess
import tempfile

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(REPO, 'include', 'alm_ocr.h')
LIB_DIR = os.path.join(REPO, 'advancedliteratemachinery_b200')


@pytest.mark.parametrize('cc,std', [('gcc', '-std=c99'), ('g++', '-std=c++17')])
def test_header_compiles_standalone(cc, std):
    if shutil.which(cc) is None:
        pytest.skip(f'{cc} not available')
    src = os.path.join(tempfile.mkdtemp(), 'h.c' if cc == 'gcc' else 'h.cpp')
    open(src, 'w').write('#include "alm_ocr.h"\nint main(void) { return ALM_OK; }\n')
    subprocess.check_call([cc, std, '-Wall', '-Werror', '-pedantic', '-fsyntax-only', '-I', os.path.dirname(HDR), src])


def test_c_host_links_and_calls_the_host_only_entry_points():
    if shutil.which('gcc') is None or not os.path.exists(os.path.join(LIB_DIR, 'libalm_ocr.so')):
        pytest.skip('gcc or libalm_ocr.so not available')
    d = tempfile.mkdtemp()
    src = os.path.join(d, 'host.c')
    open(src, 'w').write(r'''
#include <stdio.h>
#include <string.h>
#include "alm_ocr.h"
int main(void) {
  int sizes[2], hm, wm, h = 480, w = 640;
  if (alm_pre_omni_plan(&h, &w, 1, 1024, 1824, sizes, &hm, &wm) != ALM_OK) return 1;
  if (sizes[0] != 1024 || sizes[1] != 1365 || hm != 1024 || wm != 1365) return 2;
  int64_t pt[2] = {500, 250}, poly[32], rec[3] = {1033, 1034, 1099};
  float prob[3] = {0.5f, 0.25f, 0.9f};
  for (int i = 0; i < 32; ++i) poly[i] = 10 * i;
  double pts[2], polys[32], score;
  char text[32];
  if (alm_post_omni_spotting(pt, poly, rec, prob, 1, 3, 1000, 1096, 1099,
        " !\"#$%&'()*+,-./0123456789:;<=>?@ABCDEFGHIJKLMNOPQRSTUVWXYZ[\\]^_`abcdefghijklmnopqrstuvwxyz{|}~",
        480, 640, pts, polys, &score, text, sizeof text) != ALM_OK) { puts(alm_post_last_error()); return 3; }
  if (strcmp(text, "AB") != 0 || pts[0] != 320.0 || pts[1] != 120.0) return 4;
  printf("%s %.3f %s\n", text, score, alm_version());
  return 0;
}
''')
    exe = os.path.join(d, 'host')
    subprocess.check_call(['gcc', '-std=c99', '-Wall', '-Werror', '-I', os.path.dirname(HDR), src, '-o', exe,
                           '-L', LIB_DIR, '-lalm_ocr', f'-Wl,-rpath,{LIB_DIR}'])
    out = subprocess.check_output([exe], text=True)
    assert out.startswith('AB 0.375')
