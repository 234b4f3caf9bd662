//! Dumps what `filter_kmers` returns for one group of reads, in the digest format of SURVEY.md Appendix B.1, so that the
//! oracle (oracle/dbg_oracle.c) can be pinned against the real `debruijn` crate the moment a Rust toolchain is available.
//!
//! The call is the reference's, argument for argument (tzeitim/rogtk src/fracture.rs:110-116):
//!     filter_kmers::<K, _, _, _, _>(seq_tuples, &Box::new(CountFilter::new(min_coverage)), true, true, 4)
//! with the reads prepared as src/fracture.rs:218-240 does (upper-case, ACGT-only, DnaString, Exts::empty()) and K chosen as
//! src/fracture.rs:246-256 does.
//!
//! usage: dump_filter_kmers <k> <min_coverage> <reads.txt>      (one read per line; a line "-" is a null row)
//! stdout: one line per surviving k-mer, ascending:  KMER<TAB>EE<TAB>COUNT<TAB>SLOT
//!         (EE = Exts byte, two lower-case hex digits; SLOT = position in BoomHashMap2 iteration order = the reference's graph node id,
//!          i.e. the `node_rank` argument of rogtk_gpu_assemble_batch)
//! stderr: n, sha256 over the first three columns ("KMER\tEE\tCOUNT\n" per line) -- compare with tests/golden/appendix_b.json
use debruijn::dna_string::DnaString;
use debruijn::filter::{filter_kmers, CountFilter};
use debruijn::kmer::{Kmer16, Kmer32, Kmer4, Kmer64, Kmer8};
use debruijn::{Exts, Kmer};
use sha2::{Digest, Sha256};
use std::io::{BufRead, BufReader};

fn dump<K: Kmer + Send + Sync>(seqs: &[(DnaString, Exts, ())], min_coverage: usize) {
    let (valid_kmers, _all) = filter_kmers::<K, _, _, _, _>(seqs, &Box::new(CountFilter::new(min_coverage)), true, true, 4);
    let mut rows: Vec<(String, u8, u16, usize)> =
        valid_kmers.iter().enumerate().map(|(slot, (kmer, exts, count))| (kmer.to_string(), exts.val, *count, slot)).collect();
    rows.sort();
    let mut h = Sha256::new();
    for (s, e, c, slot) in &rows {
        h.update(format!("{}\t{:02x}\t{}\n", s, e, c).as_bytes());
        println!("{}\t{:02x}\t{}\t{}", s, e, c, slot);
    }
    eprintln!("n={} sha256={:x}", rows.len(), h.finalize());
}

fn main() {
    let a: Vec<String> = std::env::args().collect();
    if a.len() != 4 {
        eprintln!("usage: dump_filter_kmers <k> <min_coverage> <reads.txt>");
        std::process::exit(2);
    }
    let k: usize = a[1].parse().expect("k");
    let min_coverage: usize = a[2].parse().expect("min_coverage");
    let f = BufReader::new(std::fs::File::open(&a[3]).expect("reads file"));
    let mut seqs: Vec<(DnaString, Exts, ())> = Vec::new();
    for line in f.lines() {
        let line = line.unwrap();
        if line == "-" {
            continue; // a null row: dropped by the plugin's .flatten() (src/expressions.rs:737-740)
        }
        let s = line.trim_end_matches(&['\r', '\n'][..]).to_uppercase();
        if s.bytes().all(|b| matches!(b, b'A' | b'C' | b'G' | b'T')) {
            seqs.push((DnaString::from_dna_string(&s), Exts::empty(), ()));
        }
    }
    if k > 64 || seqs.is_empty() {
        eprintln!("n=0 (k > 64 or no valid reads: the reference returns no contigs, src/fracture.rs:212-236)");
        return;
    }
    match k {
        k if k <= 4 => dump::<Kmer4>(&seqs, min_coverage),
        k if k <= 8 => dump::<Kmer8>(&seqs, min_coverage),
        k if k <= 16 => dump::<Kmer16>(&seqs, min_coverage),
        k if k <= 32 => dump::<Kmer32>(&seqs, min_coverage),
        _ => dump::<Kmer64>(&seqs, min_coverage),
    }
}
