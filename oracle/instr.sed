# Stream edit applied to the reference qv.cc on its way into the compiler (oracle/Makefile).
# Adds three observation probes; no statement of the reference is changed or removed.
#   VG_TRACE=<file>  per read (final orientation pass): vote result + every recorded hit context
#   VG_DUMP=<file>   after the read loop: every SNP site of the dense pileup; then one line per call
# Anchors are line numbers of the surveyed revision (sha256 checked by the Makefile).
23a\
static unsigned long vg_read_no = 0; static FILE *vg_trace_f = NULL; static FILE *vg_dump_f = NULL;
753a\
	{ const char *vg_p = getenv("VG_TRACE"); if (vg_p) { vg_trace_f = fopen(vg_p, "w"); assert(vg_trace_f); } \
	  vg_p = getenv("VG_DUMP"); if (vg_p) { vg_dump_f = fopen(vg_p, "w"); assert(vg_dump_f); } }
760a\
		++vg_read_no;
1510a\
		if (vg_trace_f) { \
			fprintf(vg_trace_f, "R %lu %d %d %u %d %d %zu %zu\\n", vg_read_no - 1, (int)revcompl, (int)process_read, target_index, \
			        index_table.best ? (int)index_table.best->freq : 0, (int)index_table.ambiguous, n_ref_hits, n_snp_hits); \
			for (size_t vg_i = 0; vg_i < n_ref_hits; vg_i++) fprintf(vg_trace_f, "r %u %u %lu %u\\n", ref_hit_contexts[vg_i].position, \
			        ref_hit_contexts[vg_i].kmer_pos, (unsigned long)ref_hit_contexts[vg_i].kmer, ref_hit_contexts[vg_i].modified_pos); \
			for (size_t vg_i = 0; vg_i < n_snp_hits; vg_i++) fprintf(vg_trace_f, "s %u %u %lu %u\\n", snp_hit_contexts[vg_i].position, \
			        snp_hit_contexts[vg_i].kmer_pos, (unsigned long)snp_hit_contexts[vg_i].kmer, snp_hit_contexts[vg_i].modified_pos); \
		}
1558a\
	if (vg_trace_f) fclose(vg_trace_f); \
	if (vg_dump_f) { for (size_t vg_i = 0; vg_i < pileup_size; vg_i++) { struct packed_pileup_entry *vg_e = &pileup_table[vg_i]; \
		if (vg_e->ref != 0 || vg_e->alt != 0) fprintf(vg_dump_f, "P %zu %u %u %u %u %u %u\\n", vg_i, (unsigned)vg_e->ref, (unsigned)vg_e->alt, \
			(unsigned)vg_e->ref_cnt, (unsigned)vg_e->alt_cnt, (unsigned)vg_e->ref_freq, (unsigned)vg_e->alt_freq); } }
1596a\
			if (vg_dump_f) fprintf(vg_dump_f, "C %zu %s %zu %d %.17g\\n", i, chrlens[j].name, index, call.genotype, call.confidence);
