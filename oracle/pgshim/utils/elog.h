/* pgshim: intentionally empty stand-in for PostgreSQL's utils/elog.h (test infrastructure only). */
