/* pgshim: intentionally empty stand-in for PostgreSQL's utils/builtins.h (test infrastructure only). */
