p='vaq_b200/csrc/vaqgpu_host.cu'
s=open(p).read()
def rep(a,b,cnt=1):
    global s
    assert s.count(a)==cnt, (s.count(a), a)
    s=s.replace(a,b)

rep("""  bool cluster_windows = false;                  // the re-ordering windows follow the TI clusters (not aligned 4096-row blocks)
""","""  bool cluster_windows = false;                  // the re-ordering windows follow the TI / scan-order clusters (not aligned 4096-row blocks)
  // scan order (EA / HEAP searches of an index without TI clusters): rows [0, oc_n) grouped by a coarse clustering of
  // their leading subspaces so that a query tile can start its scan at the rows nearest to it (ensure_layout)
  int32_t oc_C = 0, oc_dims = 0;
  int64_t oc_n = 0;
  float *d_oc_centres_t = nullptr;               // [oc_dims][oc_C]
  int64_t *d_oc_start = nullptr, *d_oc_size = nullptr;
""")
rep("""  cudaFree(h->d_rowid);
  for (DevBuf *b""","""  cudaFree(h->d_rowid);
  cudaFree(h->d_oc_centres_t); cudaFree(h->d_oc_start); cudaFree(h->d_oc_size);
  for (DevBuf *b""")
old_start = s.index("// Re-orders the windows that hold rows added since the last call")
old_end = s.index("// Back to the arrival order (TI cluster ranges are defined on it)")
new_layout = open('scripts/dev/new_layout.txt').read()
s = s[:old_start] + new_layout + s[old_end:]
rep("""  h->opt_n = 0;
  h->cluster_windows = false;
  return VAQGPU_OK;
}

// Appending rows (or replacing the clusters)""","""  h->opt_n = 0;
  h->cluster_windows = false;
  clear_scan_order(h);
  return VAQGPU_OK;
}

// Appending rows (or replacing the clusters)""")
open(p,'w').write(s)
