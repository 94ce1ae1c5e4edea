/* pgshim: intentionally empty stand-in for PostgreSQL's utils/varlena.h (test infrastructure only). */
