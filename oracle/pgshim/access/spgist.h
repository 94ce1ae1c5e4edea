/* pgshim: intentionally empty stand-in for PostgreSQL's access/spgist.h (test infrastructure only). */
