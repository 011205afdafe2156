// cnf_io.h -- DIMACS CNF loading behind the cnf_io API the reference CLI uses
// (example/cnf_io/cnf_io.h:13,16 of the reference: cnf_header_read / cnf_data_read, return TRUE ON ERROR).
//
// A from-scratch single-pass loader: the file is mapped once, cut at line starts into one piece per host thread
// and scanned with a hand-written integer tokenizer; cnf_header_read caches the parse so the cnf_data_read that
// follows does not touch the file again (the reference parses the text twice, copying the rest of the line for
// every token).  ALLL_CNF_THREADS overrides the thread count (default: all hardware threads, at most 64).
//
// Dialect: 'c'/'C' comment lines anywhere; first other non-blank line is
// "p cnf V C" (case-insensitive, any blanks); clauses are integer streams terminated by 0, may span lines,
// several per line.  Deliberate deviations from the reference's hazards: a last line without '\n' IS
// parsed; tabs and '\r' separate tokens; a line starting with '%' ends the data (SATLIB trailer) instead of
// triggering an out-of-bounds write; counts that contradict the header are reported as an error.
#ifndef ALLL_B200_CNF_IO_H
#define ALLL_B200_CNF_IO_H

#include <cstdint>
#include <string>
#include <vector>

using namespace std;

// Reads V, C from the problem line and counts the non-zero literals.  Returns true on error.
bool cnf_header_read(const string &cnf_file_name, int *v_num, int *c_num, int *l_num);

// Fills l_c_num[c_num] (literals per clause) and l_val[l_num] (signed literals, clause by clause).
// Never writes outside the given sizes.  Returns true on error (unreadable file, bad token, count mismatch).
bool cnf_data_read(const string &cnf_file_name, int v_num, int c_num, int l_num, int l_c_num[], int l_val[]);

// Extension (not in the reference): the same parse delivered directly as the CSR the C ABI uploads
// (alll_upload_csr): off[m+1], lit[L] with the reference's literal encoding x>0 -> 2x-2, x<0 -> -2x-1
// (example/main.cpp:168).  Skips the int arrays and the per-clause object graph of example/main.cpp:157-178.
// *c_num_header is the clause count the problem line announces; off.size()-1 is what the file holds.
// Returns true on error (unreadable file, bad token, count mismatch, variable index beyond V).
bool cnf_read_csr(const string &cnf_file_name, int *v_num, int *c_num_header, vector<uint64_t> &off, vector<uint32_t> &lit,
                  int n_threads = 0);

#endif
