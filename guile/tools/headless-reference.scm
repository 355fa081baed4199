;;; headless-reference.scm — run the REFERENCE's own per-tick update without a window, dump the world
;;; after every tick, and time the loop.  SURVEY.md 8(f) rank 2: the true Guile CPU baseline and the
;;; golden vectors the reference does not ship.
;;;
;;;   guile -L /path/to/griddy -L /path/to/this/repo/guile \
;;;         guile/tools/headless-reference.scm /path/to/griddy/examples/triangle.scm 10000 > triangle.dump
;;;
;;; Nothing of the reference is modified or re-implemented: `advance$' and `make-++' are taken out of
;;; (griddy simulate) with module-ref (they are module-private, simulate.scm:17), and the loop below is
;;; the body of `update' (simulate.scm:144-151) minus the wall-clock bookkeeping.  The example script's
;;; final (simulate make-skeleton add-actors! ...) is intercepted to capture the two procedures instead
;;; of opening the SDL window (event-loop.scm:192-211).
;;;
;;; Output, one block per tick, actors in `get-actors' order (world.scm:24-30):
;;;   tick T N
;;;   <on-road 0|1> <container registration index> <side 0 forw|1 back> <pos-param> <agenda length>
;;;       <head sleep time or 0> <route: none|done|lane index|arrive> <max-speed> <x> <y>
;;; Numbers are printed with number->string, which round-trips doubles exactly.  tests/golden/
;;; check_guile_dump.py compares such a dump with a fixture of tests/golden/ tick by tick, bit for bit.
;;; The last line on stderr is the baseline: living-actor updates per second of the Guile loop.
;;;
;;; Not executed under Guile here (no Guile, Chickadee, pfds in the build image).  tests/test_guile_glue.py runs
;;; this very file under the fixtures' Scheme evaluator and feeds its output to check_guile_dump.py.
(use-modules (oop goops)
             (srfi srfi-1)
             (srfi srfi-26)
             (ice-9 format)
             (chickadee math vector)
             (griddy core)
             (griddy util)
             (griddy simulate))

(define args (command-line))
(unless (= (length args) 3)
  (format (current-error-port) "usage: ~a EXAMPLE.scm TICKS~%" (car args))
  (exit 2))
(define example (cadr args))
(define ticks (string->number (caddr args)))

(define sim-module (resolve-module '(griddy simulate)))
(define advance$ (module-ref sim-module 'advance$))
(define make-++ (module-ref sim-module 'make-++))

(define captured #f)
(module-set! sim-module 'simulate
             (lambda (make-skeleton add-actors! . _)
               (set! captured (cons make-skeleton add-actors!))))
(load example)
(unless captured
  (format (current-error-port) "~a did not call simulate~%" example)
  (exit 2))

(define (registration-index world)
  (let* ((items (get-static-items world))
         (n (length items))
         (table (make-hash-table)))
    (let loop ((l items) (i 0))
      (unless (null? l)
        (hashq-set! table (car l) (- n 1 i))      ; world.scm:21-22 conses
        (loop (cdr l) (1+ i))))
    table))

(define (dump tick world)
  (let ((index (registration-index world))
        (actors (get-actors world)))
    (format #t "tick ~a ~a~%" tick (length actors))
    (for-each
     (lambda (actor)
       (let* ((loc (ref actor 'location))
              (on-road (is-a? loc <location/on-road>))
              (agenda (ref actor 'agenda))
              (route (ref actor 'route))
              (p (get-pos actor)))
         (format #t "~a ~a ~a ~a ~a ~a ~a ~a ~a ~a~%"
                 (if on-road 1 0)
                 (hashq-ref index (ref loc (if on-road 'road-lane 'road-segment)))
                 (if (and (not on-road) (eq? (ref loc 'road-side-direction) 'back)) 1 0)
                 (number->string (exact->inexact (ref loc 'pos-param)))
                 (length agenda)
                 (if (and (pair? agenda) (eq? (caar agenda) 'sleep-for)) (cadar agenda) 0)
                 (cond ((eq? route 'none) "none")
                       ((null? route) "done")
                       ((eq? (caar route) 'turn-onto) (hashq-ref index (cadar route)))
                       (else "arrive"))
                 (number->string (exact->inexact (ref actor 'max-speed)))
                 (number->string (vec2-x p))
                 (number->string (vec2-y p)))))
     actors)))

(let* ((make-skeleton (car captured))
       (add-actors! (cdr captured))
       (world (make-skeleton))
       (updates 0)
       (busy 0))
  (add-actors! world)
  (dump 0 world)
  (do ((t 1 (1+ t))) ((> t ticks))
    (let* ((t0 (get-internal-real-time))
           (world++ (make-skeleton))                           ; simulate.scm:146
           (++ (make-++ world world++))                        ; simulate.scm:147
           (actors (get-actors world)))
      (for-each (cut advance$ ++ <>) actors)                   ; simulate.scm:148-150
      (set! busy (+ busy (- (get-internal-real-time) t0)))
      (set! updates (+ updates (length actors)))
      (set! world world++)                                     ; simulate.scm:151
      (dump t world)))
  (let ((seconds (/ busy internal-time-units-per-second)))
    (format (current-error-port) "~a actor-updates in ~,3f s of update time: ~,1f actor-updates/s (Guile ~a, 1 thread)~%"
            updates (exact->inexact seconds)
            (if (zero? seconds) 0.0 (exact->inexact (/ updates seconds)))
            (version))))
