-- The CPU baseline north_star names: the STOCK reference extension inside PostgreSQL, parallel scan on all host cores.
-- Not executable in this image (no PostgreSQL, no network); run it wherever a PostgreSQL >= 13 with the reference's
-- `make install` exists:
--     python tools/pg_baseline/make_reads_csv.py 1000000 /tmp/reads_1g.csv
--     psql -v cores=$(nproc) -v csv=/tmp/reads_1g.csv -f tools/pg_baseline/baseline.sql
-- k-mers/s = (rows * (1000 - 21 + 1)) / execution time of the last statement.
CREATE EXTENSION IF NOT EXISTS kmer;                       -- kmer--1.0.0.sql, unchanged
DROP TABLE IF EXISTS reads;
CREATE TABLE reads (dna dna);
\set copycmd '\\copy reads FROM ' :'csv'
:copycmd
ANALYZE reads;
SET max_parallel_workers_per_gather = :cores;
SET max_parallel_workers = :cores;
SET parallel_setup_cost = 0;
SET parallel_tuple_cost = 0;
SET min_parallel_table_scan_size = 0;
SET work_mem = '4GB';                                      -- state it with the result: HashAggregate spills beyond it
-- hash(kmer) is not marked PARALLEL SAFE (kmer--1.0.0.sql:133-136): check that a parallel plan is chosen
EXPLAIN SELECT kmer, count(*) FROM (SELECT generate_kmers(dna, 21) AS kmer FROM reads) s GROUP BY kmer;
-- configs[0] as well: SELECT ... generate_kmers(dna, 5) ... over 10 000 reads
EXPLAIN (ANALYZE, BUFFERS)
SELECT kmer, count(*) FROM (SELECT generate_kmers(dna, 21) AS kmer FROM reads) s GROUP BY kmer;
