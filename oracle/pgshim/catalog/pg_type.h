/* pgshim: intentionally empty stand-in for PostgreSQL's catalog/pg_type.h (test infrastructure only). */
